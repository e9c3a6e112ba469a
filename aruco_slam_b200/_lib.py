"""ctypes loader of the product library `csrc/libb2aruco.so` (C ABI: include/b2aruco.h).

There is no CPU fallback: if the library is missing, or no CUDA device can be
opened, the calls raise -- loudly -- instead of computing anything on the host.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("B2A_LIB", os.path.join(_HERE, "csrc", "libb2aruco.so"))        # B2A_LIB: an instrumented build (profiling only)


class B2AError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b2aruco error {code}: {msg}")
        self.code = code


class DetectorParams(C.Structure):
    """cv::aruco::DetectorParameters (cv2 4.13.0 defaults via b2a_default_detector_params)."""
    _fields_ = [
        ("adaptiveThreshWinSizeMin", C.c_int), ("adaptiveThreshWinSizeMax", C.c_int),
        ("adaptiveThreshWinSizeStep", C.c_int), ("adaptiveThreshConstant", C.c_double),
        ("minMarkerPerimeterRate", C.c_double), ("maxMarkerPerimeterRate", C.c_double),
        ("polygonalApproxAccuracyRate", C.c_double), ("minCornerDistanceRate", C.c_double),
        ("minDistanceToBorder", C.c_int), ("minMarkerDistanceRate", C.c_double),
        ("minGroupDistance", C.c_float), ("markerBorderBits", C.c_int),
        ("perspectiveRemovePixelPerCell", C.c_int), ("perspectiveRemoveIgnoredMarginPerCell", C.c_double),
        ("maxErroneousBitsInBorderRate", C.c_double), ("minOtsuStdDev", C.c_double),
        ("errorCorrectionRate", C.c_double), ("cornerRefinementMethod", C.c_int),
        ("cornerRefinementWinSize", C.c_int), ("relativeCornerRefinmentWinSize", C.c_double),
        ("cornerRefinementMaxIterations", C.c_int), ("cornerRefinementMinAccuracy", C.c_double),
        ("detectInvertedMarker", C.c_int),
        ("useAruco3Detection", C.c_int), ("minSideLengthCanonicalImg", C.c_int), ("minMarkerLengthRatioOriginalImg", C.c_float),
    ]


class CDictionary(C.Structure):
    _fields_ = [("markerSize", C.c_int), ("maxCorrectionBits", C.c_int), ("nMarkers", C.c_int),
                ("nBytes", C.c_int), ("table", C.c_void_p)]


class DetectorConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("max_width", C.c_int), ("max_height", C.c_int), ("max_batch", C.c_int),
                ("max_markers", C.c_int), ("max_candidates", C.c_int)]


class Frames(C.Structure):
    _fields_ = [("data", C.c_void_p), ("on_device", C.c_int), ("batch", C.c_int), ("width", C.c_int),
                ("height", C.c_int), ("channels", C.c_int), ("row_stride", C.c_size_t), ("frame_stride", C.c_size_t)]


class Detections(C.Structure):
    _fields_ = [("batch", C.c_int), ("max_markers", C.c_int),
                ("n_accepted", C.POINTER(C.c_int32)), ("n_rejected", C.POINTER(C.c_int32)),
                ("corners", C.POINTER(C.c_float)), ("ids", C.POINTER(C.c_int32)), ("rejected", C.POINTER(C.c_float)),
                ("rvecs", C.POINTER(C.c_double)), ("tvecs", C.POINTER(C.c_double)), ("status", C.POINTER(C.c_int32))]


class Camera(C.Structure):
    _fields_ = [("K", C.c_double * 9), ("D", C.c_double * 5), ("nD", C.c_int), ("marker_length", C.c_float)]


class Board(C.Structure):
    _fields_ = [("n_markers", C.c_int), ("ids", C.c_void_p), ("obj_points", C.c_void_p)]


class RefineParams(C.Structure):
    _fields_ = [("minRepDistance", C.c_float), ("errorCorrectionRate", C.c_float), ("checkAllOrders", C.c_int)]


class SlamParams(C.Structure):
    _fields_ = [("Q_k", C.c_double), ("R_x", C.c_double), ("R_y", C.c_double), ("R_theta", C.c_double),
                ("kl", C.c_double), ("kr", C.c_double), ("b", C.c_double), ("r2c_tx", C.c_double), ("r2c_ty", C.c_double),
                ("useful_distance_threshold", C.c_float), ("max_landmarks", C.c_int)]


class Observation(C.Structure):
    _fields_ = [("aruco_id", C.c_int32), ("aruco_index", C.c_int32), ("x", C.c_double), ("y", C.c_double),
                ("theta", C.c_double), ("cov", C.c_double * 9)]


class MapMarker(C.Structure):
    _fields_ = [("id", C.c_int32), ("length", C.c_double), ("x", C.c_double), ("y", C.c_double), ("z", C.c_double),
                ("roll", C.c_double), ("pitch", C.c_double), ("yaw", C.c_double), ("q", C.c_double * 4)]


class PoseWithCovariance(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("orientation", C.c_double * 4), ("covariance", C.c_double * 36)]


# every symbol include/b2aruco.h declares
SYMBOLS = [
    "b2a_last_error", "b2a_version", "b2a_default_detector_params", "b2a_get_predefined_dictionary",
    "b2a_detector_create", "b2a_detector_destroy", "b2a_detect", "b2a_detect_pose", "b2a_detect_pose_submit", "b2a_detect_pose_wait", "b2a_detector_set_inflight",
    "b2a_estimate_pose_single_markers", "b2a_debug_threshold", "b2a_debug_contours", "b2a_debug_candidates",
    "b2a_detector_num_scales", "b2a_detector_set_streams", "b2a_last_stage_times", "b2a_last_launch_count", "b2a_detector_stream",
    "b2a_default_slam_params", "b2a_slam_create", "b2a_slam_destroy", "b2a_slam_dim", "b2a_slam_get_state",
    "b2a_slam_set_state", "b2a_slam_add_encoder", "b2a_slam_make_observations", "b2a_slam_update", "b2a_slam_add_image", "b2a_slam_add_image_submit", "b2a_slam_add_image_wait", "b2a_slam_synchronize",
    "b2a_quaternion_from_rpy", "b2a_map_parse", "b2a_map_load", "b2a_slam_robot_pose", "b2a_slam_detected_map",
    "b2a_pack_robot_pose", "b2a_pack_map_marker", "b2a_slam_stream",
    "b2a_multi_create", "b2a_multi_destroy", "b2a_multi_num_devices", "b2a_multi_detect_pose", "b2a_draw_detected_markers", "b2a_detector_last_detections", "b2a_pack_detections",
    "b2a_default_refine_params", "b2a_refine_detected_markers", "b2a_detector_set_graph", "b2a_slam_robot_pose_submit", "b2a_slam_robot_pose_wait",
    "b2a_host_alloc", "b2a_host_free", "b2a_pack_detected_markers",
]

_lib = None


def build(force: bool = False) -> str:
    """Compile csrc/libb2aruco.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-s", "libb2aruco.so"]
    if force:
        args.insert(1, "-B")
    subprocess.check_call(args)
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise B2AError(-1, f"{SO_PATH} is not built (run __graft_entry__.build() or make -C aruco_slam_b200/csrc); "
                               "there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        L.b2a_last_error.restype = C.c_char_p
        L.b2a_version.restype = C.c_char_p
        L.b2a_detector_stream.restype = C.c_void_p
        for name in ("b2a_detector_destroy", "b2a_detector_num_scales", "b2a_last_launch_count", "b2a_detector_stream",
                     "b2a_slam_destroy", "b2a_slam_dim", "b2a_slam_synchronize"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.b2a_quaternion_from_rpy.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.b2a_quaternion_from_rpy.restype = None
        L.b2a_map_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
        L.b2a_map_load.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p]
        L.b2a_slam_robot_pose.argtypes = [C.c_void_p, C.c_void_p]
        L.b2a_slam_detected_map.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
        L.b2a_detector_set_streams.argtypes = [C.c_void_p, C.c_int]
        L.b2a_detector_set_inflight.argtypes = [C.c_void_p, C.c_int]
        L.b2a_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_detect_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_multi_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_multi_destroy.argtypes = [C.c_void_p]
        L.b2a_detector_last_detections.argtypes = [C.c_void_p, C.c_void_p]
        L.b2a_pack_detections.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.b2a_draw_detected_markers.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.b2a_multi_detect_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_slam_add_image_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_slam_add_image_wait.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b2a_slam_stream.argtypes = [C.c_void_p]
        L.b2a_slam_stream.restype = C.c_void_p
        L.b2a_detect_pose_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_detect_pose_wait.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.b2a_estimate_pose_single_markers.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_default_refine_params.argtypes = [C.c_void_p]
        L.b2a_default_refine_params.restype = None
        L.b2a_refine_detected_markers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_debug_threshold.argtypes = [C.c_void_p] * 4
        L.b2a_debug_contours.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.b2a_debug_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b2a_last_stage_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.b2a_slam_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.b2a_slam_get_state.argtypes = [C.c_void_p] * 4
        L.b2a_slam_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_slam_add_encoder.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.b2a_slam_make_observations.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
        L.b2a_slam_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.b2a_slam_add_image.argtypes = [C.c_void_p] * 4
        L.b2a_pack_detected_markers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.b2a_host_alloc.argtypes = [C.c_size_t, C.c_int, C.c_void_p]
        L.b2a_host_free.argtypes = [C.c_void_p]
        L.b2a_host_free.restype = None
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise B2AError(rc, lib().b2a_last_error().decode())


class HostBuffer:
    """b2a_host_alloc'd pinned staging memory as a numpy array (`.array`); freed with the object."""

    def __init__(self, shape, dtype=np.uint8, write_combined=False):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(lib().b2a_host_alloc(n, 1 if write_combined else 0, C.byref(p)))
        self._p = p
        self.array = np.ctypeslib.as_array((C.c_uint8 * n).from_address(p.value)).view(dtype).reshape(shape)

    def close(self):
        if self._p is not None and self._p.value:
            self.array = None
            lib().b2a_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
