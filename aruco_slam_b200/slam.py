"""Host-side mirror of the reference's `ArucoSlam` class (include/aruco_slam/aruco_slam.h:101-193):
addEncoder / addImage / setCameraParameters and state accessors, with the detection, pose,
observation mapping and EKF arithmetic running on the GPU behind include/b2aruco.h.
No ROS types: poses and covariances are plain arrays; `dt` is explicit (the reference reads
ros::Time::now(), src/aruco_slam.cpp:26,31-32)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .aruco import ArucoDetector, DetectorParameters, _camera
from .dictionaries import getPredefinedDictionary


def slam_params(**kw) -> _lib.SlamParams:
    """defaults: reference parameters.yaml:5-13 and aruco_slam.h:58"""
    p = _lib.SlamParams()
    _lib.lib().b2a_default_slam_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class ArucoSlam:
    def __init__(self, markers_dictionary: int = 16, marker_length: float = 0.27, *, image_shape=(480, 640), device: int = 0,
                 detector_parameters=None, **slam_kw):
        self.marker_length = float(marker_length)
        self.params = slam_params(**slam_kw)
        h = C.c_void_p()
        _lib.check(_lib.lib().b2a_slam_create(int(device), C.byref(self.params), C.byref(h)))
        self._h = h
        self.detector = ArucoDetector(getPredefinedDictionary(markers_dictionary), detector_parameters or DetectorParameters(),
                                      max_shape=image_shape, max_batch=1, device=device)
        self._cam = None

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().b2a_slam_destroy(self._h)
            self._h = None
        if getattr(self, "detector", None):
            self.detector.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setCameraParameters(self, camera_matrix, dist_coeffs):
        """aruco_slam.h:129-133"""
        self._cam = _camera(camera_matrix, dist_coeffs, self.marker_length)

    def addEncoder(self, wl: float, wr: float, dt: float):
        """src/aruco_slam.cpp:21-74.  As in the reference (:24-29) the first call on a fresh filter only marks it initialised
        (there: latches the clock) whatever its arguments; dt=None reads as 0."""
        _lib.check(_lib.lib().b2a_slam_add_encoder(self._h, float(wl), float(wr), 0.0 if dt is None else float(dt)))

    def addImage(self, image):
        """src/aruco_slam.cpp:76-263"""
        if self._cam is None:
            raise _lib.B2AError(1, "setCameraParameters first")
        fr, keep = ArucoDetector._frames_host(image)
        _lib.check(_lib.lib().b2a_slam_add_image(self._h, self.detector._h, C.byref(fr), C.byref(self._cam)))
        self._last_image = keep[0]

    def getMarkedImg(self):
        """the last frame with cv::aruco::drawDetectedMarkers applied (aruco_slam.cpp:318-319, aruco_slam.h:152); drawn on demand from
        the detections the last addImage left in the detector's host arrays"""
        img = getattr(self, "_last_image", None)
        if img is None:
            return None
        try:
            r = self.detector.last_detections()
        except _lib.B2AError:
            return None                                        # addImage before the first encoder message does not look at the frame (:84-85)
        return self.detector.drawDetectedMarkers(img.copy(), r.corners[0], r.ids[0]) if len(r.ids[0]) else img.copy()

    def toRosDetectedMarkers(self, r2c_q=None, r2c_t=None):
        """the cubes of the last frame's markers inside the useful range, in the robot frame (detected_markers_, aruco_slam.cpp:324-347);
        r2c_q / r2c_t = rotation (x, y, z, w) and translation of transformStamped_r2c (default: identity, (r2c_tx, r2c_ty, 0))"""
        from . import formats
        try:
            r = self.detector.last_detections()
        except _lib.B2AError:
            return []
        if r.rvecs is None or not len(r.ids[0]):
            return []
        t = (self.params.r2c_tx, self.params.r2c_ty, 0.0) if r2c_t is None else r2c_t
        return formats.detected_markers(r.ids[0], r.rvecs[0], r.tvecs[0], self.marker_length, self.params.useful_distance_threshold, r2c_q, t)

    def addImageFrames(self, frames):
        """addImage on a one-frame b2a_frames descriptor (host or device memory, see ArucoDetector.frames_device)"""
        if self._cam is None:
            raise _lib.B2AError(1, "setCameraParameters first")
        _lib.check(_lib.lib().b2a_slam_add_image(self._h, self.detector._h, C.byref(frames), C.byref(self._cam)))

    def submitImageFrames(self, frames) -> int:
        """the detection half of addImage, enqueued (b2a_slam_add_image_submit); returns the ticket for waitImage"""
        if self._cam is None:
            raise _lib.B2AError(1, "setCameraParameters first")
        t = C.c_int(-1)
        _lib.check(_lib.lib().b2a_slam_add_image_submit(self._h, self.detector._h, C.byref(frames), C.byref(self._cam), C.byref(t)))
        return t.value

    def waitImage(self, ticket: int):
        """the filter half of addImage for a submitted frame (b2a_slam_add_image_wait)"""
        _lib.check(_lib.lib().b2a_slam_add_image_wait(self._h, self.detector._h, int(ticket)))

    def make_observations(self, corners, ids, rvecs, tvecs):
        c = np.ascontiguousarray(corners, np.float32).reshape(-1, 8)
        n = len(c)
        ids = np.ascontiguousarray(ids, np.int32).reshape(-1)
        out = (_lib.Observation * max(n, 1))()
        k = C.c_int(0)
        _lib.check(_lib.lib().b2a_slam_make_observations(
            self._h, c.ctypes.data, ids.ctypes.data, np.ascontiguousarray(rvecs, np.float64).ctypes.data,
            np.ascontiguousarray(tvecs, np.float64).ctypes.data, n, C.byref(self._cam), out, C.byref(k)))
        return [out[i] for i in range(k.value)]

    def update(self, observations):
        n = len(observations)
        arr = (_lib.Observation * max(n, 1))(*observations)
        _lib.check(_lib.lib().b2a_slam_update(self._h, arr, n))

    @staticmethod
    def pack_observations(observations):
        """the C array b2a_slam_update takes (build it once when the same frame is replayed)"""
        n = len(observations)
        return (_lib.Observation * max(n, 1))(*observations), n

    def update_packed(self, packed):
        _lib.check(_lib.lib().b2a_slam_update(self._h, packed[0], packed[1]))

    def synchronize(self):
        """wait for the EKF kernels enqueued so far"""
        _lib.check(_lib.lib().b2a_slam_synchronize(self._h))

    @property
    def stream(self) -> int:
        """the cudaStream_t the EKF kernels run on"""
        return int(_lib.lib().b2a_slam_stream(self._h) or 0)

    @property
    def dim(self) -> int:
        return _lib.lib().b2a_slam_dim(self._h)

    def get_state(self):
        N = self.dim
        mu = np.zeros(N)
        sg = np.zeros((N, N))
        ids = np.zeros(max((N - 3) // 3, 1), np.int32)
        _lib.check(_lib.lib().b2a_slam_get_state(self._h, mu.ctypes.data, sg.ctypes.data, ids.ctypes.data))
        return mu, sg, ids[:(N - 3) // 3]

    def set_state(self, mu, sigma, ids):
        mu = np.ascontiguousarray(mu, np.float64)
        sigma = np.ascontiguousarray(sigma, np.float64)
        ids = np.ascontiguousarray(ids, np.int32)
        _lib.check(_lib.lib().b2a_slam_set_state(self._h, len(mu), mu.ctypes.data, sigma.ctypes.data, ids.ctypes.data if len(ids) else None))

    def robot_pose(self):
        """toRosPose() without ROS (src/aruco_slam.cpp:378-410): (x, y, theta) and the 3x3 covariance block."""
        mu, sg, _ = self.get_state()
        return mu[:3].copy(), sg[:3, :3].copy()
