"""Host-side mirror of the call surface the reference uses (src/aruco_slam.cpp:11-12, 313-314):

    dictionary = getPredefinedDictionary(id)
    corners, ids, rejected = detectMarkers(image, dictionary, parameters)
    rvecs, tvecs = estimatePoseSingleMarkers(corners, markerLength, cameraMatrix, distCoeffs)

with the same argument meaning and result shapes as the `cv2.aruco` Python binding
(corners: tuple of (1,4,2) float32; ids: (N,1) int32 or None; rvecs/tvecs: (N,1,3) float64),
plus batched entry points (`ArucoDetector.detect_batch` / `detect_pose_batch`) that the
benchmark uses.  All compute happens in the CUDA library behind include/b2aruco.h.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import B2AError, DetectorParams
from .dictionaries import Dictionary, getPredefinedDictionary  # noqa: F401  (re-export)

CORNER_REFINE_NONE, CORNER_REFINE_SUBPIX = 0, 1


def DetectorParameters(**overrides) -> DetectorParams:
    """cv::aruco::DetectorParameters with the library's defaults."""
    p = DetectorParams()
    _lib.lib().b2a_default_detector_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _cdict(dic: Dictionary):
    t = np.ascontiguousarray(dic.table, np.uint8)
    cd = _lib.CDictionary(dic.marker_size, dic.max_correction_bits, dic.n_markers, dic.n_bytes, t.ctypes.data)
    return cd, t


def library_dictionary(dict_id: int) -> Dictionary:
    """The table compiled into the library (b2a_get_predefined_dictionary)."""
    cd = _lib.CDictionary()
    _lib.check(_lib.lib().b2a_get_predefined_dictionary(int(dict_id), C.byref(cd)))
    n = cd.nMarkers * 4 * cd.nBytes
    buf = (C.c_uint8 * n).from_address(cd.table)
    table = np.frombuffer(buf, np.uint8).reshape(cd.nMarkers, 4, cd.nBytes).copy()
    return Dictionary(dict_id, "lib", cd.markerSize, cd.maxCorrectionBits, table)


def _camera(K, D, marker_length) -> _lib.Camera:
    cam = _lib.Camera()
    K = np.asarray(K, np.float64).reshape(9)
    for i in range(9):
        cam.K[i] = K[i]
    D = np.zeros(0) if D is None else np.asarray(D, np.float64).reshape(-1)
    if len(D) not in (0, 4, 5):
        raise B2AError(1, "distCoeffs must have 0, 4 or 5 entries")
    for i in range(len(D)):
        cam.D[i] = D[i]
    cam.nD = len(D)
    cam.marker_length = float(marker_length)
    return cam


@dataclass
class BatchDetections:
    """Per-frame lists in the reference's output order."""
    corners: list      # [frame] -> (n,4,2) float32
    ids: list          # [frame] -> (n,) int32
    rejected: list     # [frame] -> (m,4,2) float32
    rvecs: list | None = None   # [frame] -> (n,3) float64
    tvecs: list | None = None


class Board:
    """cv2.aruco.Board: object points (n, 4, 3) of every marker's corners and the markers' ids"""

    def __init__(self, objPoints, ids):
        self.objPoints = np.ascontiguousarray(np.asarray(objPoints, np.float32).reshape(-1, 4, 3))
        self.ids = np.ascontiguousarray(np.asarray(ids, np.int32).ravel())
        if len(self.ids) != len(self.objPoints):
            raise B2AError(1, "objPoints and ids differ in length")

    def getObjPoints(self):
        return self.objPoints

    def getIds(self):
        return self.ids


class RefineParameters:
    """cv2.aruco.RefineParameters"""

    def __init__(self, minRepDistance=10.0, errorCorrectionRate=3.0, checkAllOrders=True):
        self.minRepDistance, self.errorCorrectionRate, self.checkAllOrders = float(minRepDistance), float(errorCorrectionRate), bool(checkAllOrders)


class ArucoDetector:
    """One detector handle (one GPU, one stream).  `max_shape` = (H, W) of the largest frame."""

    def __init__(self, dictionary: Dictionary, parameters: DetectorParams | None = None, *, max_shape=(1080, 1920),
                 max_batch: int = 1, device: int = 0, max_markers: int = 256, max_candidates: int = 2048):
        self.dictionary = dictionary
        self.parameters = parameters or DetectorParameters()
        cfg = _lib.DetectorConfig(int(device), int(max_shape[1]), int(max_shape[0]), int(max_batch), int(max_markers), int(max_candidates))
        cd, self._keep = _cdict(dictionary)
        h = C.c_void_p()
        _lib.check(_lib.lib().b2a_detector_create(C.byref(cfg), C.byref(cd), C.byref(self.parameters), C.byref(h)))
        self._h = h
        self.max_markers = max_markers
        self.max_batch = max_batch
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().b2a_detector_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- frames descriptor -------------------------------------------------------------------
    @staticmethod
    def _frames_host(images: np.ndarray):
        a = np.asarray(images)
        if a.dtype != np.uint8:
            raise B2AError(1, "image must be uint8 (CV_8UC1 or CV_8UC3)")
        if a.ndim == 2:
            a = a[None]
        elif a.ndim == 3 and a.shape[2] == 3 and a.shape[0] != 3:
            a = a[None]                 # one (H,W,3) colour frame
        if a.ndim == 3:
            ch = 1
        elif a.ndim == 4 and a.shape[3] in (1, 3):
            ch = a.shape[3]
        else:
            raise B2AError(1, "expected (H,W), (H,W,3), (B,H,W) or (B,H,W,3) uint8")
        a = np.ascontiguousarray(a)
        B, H, W = a.shape[:3]
        if a.size == 0:
            raise B2AError(1, "empty image")
        return _lib.Frames(a.ctypes.data, 0, B, W, H, ch, 0, 0), a

    @staticmethod
    def frames_device(ptr: int, batch: int, height: int, width: int, channels: int = 1, row_stride: int = 0, frame_stride: int = 0):
        """Frames already resident in device memory (e.g. a torch uint8 CUDA tensor's data_ptr())."""
        return _lib.Frames(int(ptr), 1, batch, width, height, channels, row_stride, frame_stride)

    # ---- batched API -------------------------------------------------------------------------
    def _collect(self, det, pose) -> BatchDetections:
        B, K = det.batch, det.max_markers
        na = np.ctypeslib.as_array(det.n_accepted, (B,))
        nr = np.ctypeslib.as_array(det.n_rejected, (B,))
        c = np.ctypeslib.as_array(det.corners, (B, K, 4, 2))
        i = np.ctypeslib.as_array(det.ids, (B, K))
        r = np.ctypeslib.as_array(det.rejected, (B, K, 4, 2))
        out = BatchDetections([c[b, :na[b]].copy() for b in range(B)], [i[b, :na[b]].copy() for b in range(B)],
                              [r[b, :nr[b]].copy() for b in range(B)])
        if pose:
            rv = np.ctypeslib.as_array(det.rvecs, (B, K, 3))
            tv = np.ctypeslib.as_array(det.tvecs, (B, K, 3))
            out.rvecs = [rv[b, :na[b]].copy() for b in range(B)]
            out.tvecs = [tv[b, :na[b]].copy() for b in range(B)]
        return out

    def detect_raw(self, frames, camera=None):
        """Run the pipeline; returns the C struct (pinned host buffers valid until the next call)."""
        det = _lib.Detections()
        if camera is None:
            _lib.check(_lib.lib().b2a_detect(self._h, C.byref(frames), C.byref(det)))
        else:
            _lib.check(_lib.lib().b2a_detect_pose(self._h, C.byref(frames), C.byref(camera), C.byref(det)))
        return det

    def submit_raw(self, frames, camera=None) -> int:
        """enqueue one batch (H2D copies + kernels) and return its ticket; up to two batches in flight per handle"""
        t = C.c_int(-1)
        _lib.check(_lib.lib().b2a_detect_pose_submit(self._h, C.byref(frames), C.byref(camera) if camera is not None else None, C.byref(t)))
        return t.value

    def wait_raw(self, ticket: int):
        """block until the batch of `ticket` is complete; the C struct's pinned buffers stay valid until the second next submit"""
        det = _lib.Detections()
        _lib.check(_lib.lib().b2a_detect_pose_wait(self._h, int(ticket), C.byref(det)))
        return det

    def detect_pose_stream(self, batches, marker_length, K, D):
        """generator over an iterable of (B, H, W) host batches: batch k+1 is submitted before batch k is waited for, so its
        PCIe copy runs under batch k's kernels (one host thread, one handle)"""
        cam = _camera(K, D, marker_length)
        pending = None
        for images in batches:
            fr, keep = self._frames_host(images) if not isinstance(images, _lib.Frames) else (images, None)
            t = self.submit_raw(fr, cam)
            if pending is not None:
                yield self._collect(self.wait_raw(pending[0]), True)
            pending = (t, fr, keep)
        if pending is not None:
            yield self._collect(self.wait_raw(pending[0]), True)

    def detect_batch(self, images) -> BatchDetections:
        fr, keep = self._frames_host(images) if not isinstance(images, _lib.Frames) else (images, None)
        return self._collect(self.detect_raw(fr), False)

    def detect_pose_batch(self, images, marker_length, K, D) -> BatchDetections:
        fr, keep = self._frames_host(images) if not isinstance(images, _lib.Frames) else (images, None)
        return self._collect(self.detect_raw(fr, _camera(K, D, marker_length)), True)

    def refineDetectedMarkers(self, image, board, detectedCorners, detectedIds, rejectedCorners, cameraMatrix=None, distCoeffs=None,
                              refineParams=None):
        """cv2.aruco.ArucoDetector.refineDetectedMarkers: -> (corners, ids, rejected, recoveredIdxs) shaped like cv2's
        (recoveredIdxs None when nothing was recovered, as cv2 leaves its output untouched then)"""
        fr, keep = self._frames_host(image)
        c = np.array([np.asarray(q, np.float32).reshape(4, 2) for q in detectedCorners], np.float32).reshape(-1, 4, 2)
        i = np.zeros(0, np.int32) if detectedIds is None else np.asarray(detectedIds, np.int32).ravel()
        r = np.array([np.asarray(q, np.float32).reshape(4, 2) for q in rejectedCorners], np.float32).reshape(-1, 4, 2)
        nd, nr = len(i), len(r)
        cap = nd + nr + 1
        cbuf = np.zeros((cap, 4, 2), np.float32); cbuf[:nd] = c
        ibuf = np.zeros(cap, np.int32); ibuf[:nd] = i
        rbuf = np.ascontiguousarray(r) if nr else np.zeros((1, 4, 2), np.float32)
        rec = np.zeros(max(nr, 1), np.int32)
        n_d, n_r, n_rec = C.c_int(nd), C.c_int(nr), C.c_int(0)
        cam = _camera(cameraMatrix, distCoeffs if distCoeffs is not None else np.zeros(5), 1.0) if cameraMatrix is not None else None
        rp = refineParams or RefineParameters()
        prm = _lib.RefineParams(float(rp.minRepDistance), float(rp.errorCorrectionRate), int(bool(rp.checkAllOrders)))
        bd = _lib.Board(len(board.ids), board.ids.ctypes.data, board.objPoints.ctypes.data)
        _lib.check(_lib.lib().b2a_refine_detected_markers(self._h, C.byref(fr), C.byref(bd), cbuf.ctypes.data, ibuf.ctypes.data, C.byref(n_d), cap,
                                                          rbuf.ctypes.data, C.byref(n_r), C.byref(cam) if cam is not None else None, C.byref(prm),
                                                          rec.ctypes.data, C.byref(n_rec)))
        corners = tuple(q.reshape(1, 4, 2).copy() for q in cbuf[:n_d.value])
        ids = ibuf[:n_d.value].reshape(-1, 1).copy() if n_d.value else None
        rejected = tuple(q.reshape(1, 4, 2).copy() for q in rbuf[:n_r.value]) if nr else ()
        return corners, ids, rejected, (rec[:n_rec.value].reshape(-1, 1).copy() if n_rec.value else None)

    # ---- cv2-shaped single-image API -----------------------------------------------------------
    def detectMarkers(self, image):
        a = np.asarray(image)
        if a.ndim not in (2, 3):
            raise B2AError(1, "detectMarkers takes one image")
        r = self.detect_batch(a if a.ndim == 2 or a.shape[2] == 3 else a[..., 0])
        corners = tuple(c.reshape(1, 4, 2) for c in r.corners[0])
        ids = r.ids[0].reshape(-1, 1) if len(r.ids[0]) else None
        rejected = tuple(c.reshape(1, 4, 2) for c in r.rejected[0])
        return corners, ids, rejected

    def estimatePoseSingleMarkers(self, corners, markerLength, cameraMatrix, distCoeffs):
        c = np.ascontiguousarray(np.asarray(corners, np.float32).reshape(-1, 8))
        n = len(c)
        if not markerLength > 0:
            raise B2AError(1, "markerLength must be > 0")
        rv = np.zeros((n, 3))
        tv = np.zeros((n, 3))
        cam = _camera(cameraMatrix, distCoeffs, markerLength)
        _lib.check(_lib.lib().b2a_estimate_pose_single_markers(self._h, c.ctypes.data, n, C.byref(cam), rv.ctypes.data, tv.ctypes.data))
        return rv.reshape(n, 1, 3), tv.reshape(n, 1, 3)

    def last_detections(self) -> BatchDetections:
        """the results of the handle's last completed call again (b2a_detector_last_detections)"""
        det = _lib.Detections()
        _lib.check(_lib.lib().b2a_detector_last_detections(self._h, C.byref(det)))
        return self._collect(det, bool(det.rvecs))

    def drawDetectedMarkers(self, image, corners, ids=None, borderColor=(0, 255, 0)):
        """cv::aruco::drawDetectedMarkers: draws into `image` ((H,W) or (H,W,3) uint8, contiguous) in place and returns it"""
        img = np.asarray(image)
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or (img.ndim == 3 and img.shape[2] != 3) or not img.flags.c_contiguous:
            raise B2AError(1, "drawDetectedMarkers takes a contiguous 8-bit image with 1 or 3 channels")
        c = np.ascontiguousarray(np.asarray(corners, np.float32).reshape(-1, 8))
        n = len(c)
        ida = None if ids is None else np.ascontiguousarray(np.asarray(ids, np.int32).reshape(-1))
        if ida is not None and len(ida) != n:
            raise B2AError(1, "ids and corners differ in length")
        col = (C.c_uint8 * 3)(*[int(v) for v in borderColor[:3]])
        _lib.check(_lib.lib().b2a_draw_detected_markers(self._h, img.ctypes.data, img.shape[1], img.shape[0], 1 if img.ndim == 2 else 3, 0,
                                                        c.ctypes.data, ida.ctypes.data if ida is not None else None, n, col))
        return img

    # ---- stage taps (parity tests) ---------------------------------------------------------------
    @property
    def num_scales(self) -> int:
        return _lib.lib().b2a_detector_num_scales(self._h)

    def debug_threshold(self, images):
        fr, keep = self._frames_host(images)
        gray = np.zeros((fr.batch, fr.height, fr.width), np.uint8)
        masks = np.zeros((fr.batch, self.num_scales, fr.height, fr.width), np.uint8)
        _lib.check(_lib.lib().b2a_debug_threshold(self._h, C.byref(fr), gray.ctypes.data, masks.ctypes.data))
        return gray, masks

    def debug_contours(self, images, cap=4096, pts_cap=None):
        fr, keep = self._frames_host(images)
        nS = self.num_scales
        pts_cap = pts_cap or (fr.width * fr.height) // 4
        counts = np.zeros((fr.batch, nS), np.int32)
        kept = np.zeros((fr.batch, nS), np.int32)
        lens = np.zeros((fr.batch, nS, cap), np.int32)
        pts = np.zeros((fr.batch, nS, pts_cap, 2), np.int16)
        _lib.check(_lib.lib().b2a_debug_contours(self._h, C.byref(fr), counts.ctypes.data, kept.ctypes.data, lens.ctypes.data, cap,
                                                 pts.ctypes.data, pts_cap))
        return counts, kept, lens, pts

    def debug_candidates(self, images, cap=2048):
        fr, keep = self._frames_host(images)
        n = np.zeros(fr.batch, np.int32)
        q = np.zeros((fr.batch, cap, 4, 2), np.float32)
        _lib.check(_lib.lib().b2a_debug_candidates(self._h, C.byref(fr), n.ctypes.data, q.ctypes.data, cap))
        return [q[b, :n[b]].copy() for b in range(fr.batch)]

    def last_stage_times(self) -> dict:
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        k = _lib.lib().b2a_last_stage_times(self._h, names, ms, 32)
        return {names[i].decode(): float(ms[i]) for i in range(k)}

    def set_streams(self, n: int):
        """number of concurrent sub-batches (CUDA streams) a call is cut into; 1 = serial stages"""
        _lib.check(_lib.lib().b2a_detector_set_streams(self._h, int(n)))

    def set_graph(self, on: bool):
        """CUDA-graph replay of small calls (default on); off = plain launches, e.g. to read per-stage times"""
        _lib.check(_lib.lib().b2a_detector_set_graph(self._h, int(bool(on))))

    def set_inflight(self, n: int):
        """batches that submit / wait keep in flight on this handle (1 .. 4 contexts, default 2)"""
        _lib.check(_lib.lib().b2a_detector_set_inflight(self._h, int(n)))

    def last_launch_count(self) -> int:
        return _lib.lib().b2a_last_launch_count(self._h)

    @property
    def stream(self) -> int:
        return int(_lib.lib().b2a_detector_stream(self._h) or 0)


def pack_detections(det, out: np.ndarray) -> int:
    """b2a_pack_detections: the compact records of a C result struct into the uint8 array `out`; returns the bytes written"""
    n = C.c_size_t(0)
    _lib.check(_lib.lib().b2a_pack_detections(C.byref(det), out.ctypes.data, out.nbytes, C.byref(n)))
    return int(n.value)


def unpack_detections(buf, with_size: bool = False):
    """the inverse of pack_detections (host side of a gather); with_size: also the number of bytes the record occupies"""
    b = np.frombuffer(memoryview(buf), np.uint8)
    magic, B, pose, K = np.frombuffer(b[:16].tobytes(), np.int32)
    if magic != 0x42324144:
        raise B2AError(1, "not a packed detection record")
    o = 16
    out = BatchDetections([], [], [], [] if pose else None, [] if pose else None)
    for _ in range(B):
        na, nr = np.frombuffer(b[o:o + 8].tobytes(), np.int32)
        o += 8
        out.ids.append(np.frombuffer(b[o:o + 4 * na].tobytes(), np.int32)); o += 4 * na
        out.corners.append(np.frombuffer(b[o:o + 32 * na].tobytes(), np.float32).reshape(na, 4, 2)); o += 32 * na
        if pose:
            out.rvecs.append(np.frombuffer(b[o:o + 24 * na].tobytes(), np.float64).reshape(na, 3)); o += 24 * na
            out.tvecs.append(np.frombuffer(b[o:o + 24 * na].tobytes(), np.float64).reshape(na, 3)); o += 24 * na
        out.rejected.append(np.frombuffer(b[o:o + 32 * nr].tobytes(), np.float32).reshape(nr, 4, 2)); o += 32 * nr
    return (out, o) if with_size else out


class MultiDetector:
    """Several GPUs of one box from one process: a batch of host frames is cut into contiguous blocks, one per device, each through
    its own handle and host thread (inside the library); the detections come back gathered in frame order.  Frames are independent,
    so there is no collective.  `devices` may list a device more than once (two handles on one GPU)."""

    def __init__(self, dictionary: Dictionary, parameters: DetectorParams | None = None, *, devices=(0,), max_shape=(1080, 1920),
                 max_batch: int = 1, max_markers: int = 256, max_candidates: int = 2048):
        prm = parameters or DetectorParameters()
        self._dict, self._keep = _cdict(dictionary)
        cfg = _lib.DetectorConfig(0, int(max_shape[1]), int(max_shape[0]), int(max_batch), int(max_markers), int(max_candidates))
        dev = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        _lib.check(_lib.lib().b2a_multi_create(dev, len(devices), C.byref(cfg), C.byref(self._dict), C.byref(prm), C.byref(h)))
        self._h = h
        self.devices = tuple(devices)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().b2a_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def detect_pose_batch(self, images, marker_length=None, K=None, D=None) -> BatchDetections:
        fr, keep = ArucoDetector._frames_host(images)
        det = _lib.Detections()
        cam = _camera(K, D, marker_length) if K is not None else None
        _lib.check(_lib.lib().b2a_multi_detect_pose(self._h, C.byref(fr), C.byref(cam) if cam is not None else None, C.byref(det)))
        return ArucoDetector._collect(None, det, cam is not None)


# ---- free functions shaped like the legacy cv::aruco API the reference calls -------------------------
_detectors: dict = {}


def _cached_detector(dictionary, parameters, shape, device=0):
    key = (dictionary.dict_id, dictionary.table.tobytes()[:64], bytes(parameters), shape, device)
    d = _detectors.get(key)
    if d is None:
        if len(_detectors) > 8:
            _detectors.pop(next(iter(_detectors))).close()
        d = _detectors[key] = ArucoDetector(dictionary, parameters, max_shape=shape, max_batch=1, device=device)
    return d


def detectMarkers(image, dictionary: Dictionary, parameters: DetectorParams | None = None):
    """cv::aruco::detectMarkers(image, dictionary, corners, ids, parameters, rejectedImgPoints)."""
    a = np.asarray(image)
    if a.size == 0:
        raise B2AError(1, "empty image")
    p = parameters or DetectorParameters()
    return _cached_detector(dictionary, p, a.shape[:2]).detectMarkers(a)


def estimatePoseSingleMarkers(corners, markerLength, cameraMatrix, distCoeffs, dictionary: Dictionary | None = None):
    """cv::aruco::estimatePoseSingleMarkers -> (rvecs (N,1,3), tvecs (N,1,3))."""
    dic = dictionary or getPredefinedDictionary(0)
    return _cached_detector(dic, DetectorParameters(), (64, 64)).estimatePoseSingleMarkers(corners, markerLength, cameraMatrix, distCoeffs)
