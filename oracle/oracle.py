"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module (see oracle/oracle.h).  The product
package `aruco_slam_b200` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """make liboracle.so when it is missing or older than its sources.  Several processes may call this at once (every rank of a
    torchrun bench checks its frames): the build runs under a file lock, into a temporary name, and is renamed into place."""
    srcs = [os.path.join(_HERE, f) for f in ("orc_detect.c", "orc_pyramid.c", "orc_pose.c", "orc_ekf.c", "oracle.h")]

    def stale():
        return force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)

    if stale():
        import fcntl
        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if stale():
                    tmp = _SO + ".tmp.%d" % os.getpid()
                    subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so", "OUT=" + tmp])
                    os.replace(tmp, _SO)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


class Params(C.Structure):
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


class Dict(C.Structure):
    _fields_ = [("markerSize", C.c_int), ("maxCorrectionBits", C.c_int), ("nMarkers", C.c_int),
                ("nBytes", C.c_int), ("table", C.c_void_p)]


class Detections(C.Structure):
    _fields_ = [("n_acc", C.c_int), ("n_rej", C.c_int), ("corners", C.POINTER(C.c_float)),
                ("ids", C.POINTER(C.c_int32)), ("rejected", C.POINTER(C.c_float)),
                ("n_cand", C.c_int), ("cand", C.POINTER(C.c_float)), ("cand_len", C.POINTER(C.c_int32)),
                ("n_sel", C.c_int), ("sel", C.POINTER(C.c_float)), ("sel_info", C.POINTER(C.c_int32)),
                ("n_scales", C.c_int), ("n_contours", C.POINTER(C.c_int32))]


class SlamParams(C.Structure):
    _fields_ = [("Q_k", C.c_double), ("R_x", C.c_double), ("R_y", C.c_double), ("R_theta", C.c_double),
                ("kl", C.c_double), ("kr", C.c_double), ("b", C.c_double), ("marker_length", C.c_double),
                ("r2c_tx", C.c_double), ("r2c_ty", C.c_double), ("useful_distance_threshold", C.c_float)]


class Observation(C.Structure):
    _fields_ = [("aruco_id", C.c_int), ("aruco_index", C.c_int), ("x", C.c_double), ("y", C.c_double),
                ("theta", C.c_double), ("cov", C.c_double * 9)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_ekf_create.restype = C.c_void_p
        _lib.orc_ekf_dim.argtypes = [C.c_void_p]
        _lib.orc_ekf_destroy.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def default_params(**kw) -> Params:
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def make_dict(dic) -> Dict:
    """dic: aruco_slam_b200.dictionaries.Dictionary (plain data)."""
    t = np.ascontiguousarray(dic.table)
    d = Dict(dic.marker_size, dic.max_correction_bits, dic.n_markers, dic.n_bytes, t.ctypes.data)
    d._keep = t
    return d


def bgr2gray(bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    H, W, _ = bgr.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_bgr2gray(_p(bgr), W, H, _p(out))
    return out


def pyr_down(gray):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    out = np.zeros(((H + 1) // 2, (W + 1) // 2), np.uint8)
    lib().orc_pyr_down(_p(gray), W, H, _p(out))
    return out


def resize_linear(gray, dW, dH):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    out = np.zeros((dH, dW), np.uint8)
    lib().orc_resize_linear(_p(gray), W, H, _p(out), int(dW), int(dH))
    return out


def adaptive_threshold(gray, k, Cconst=7.0):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_adaptive_threshold(_p(gray), W, H, int(k), C.c_double(Cconst), _p(out))
    return out


def find_contours(mask):
    """list of (n,2) int32 arrays in cv2.findContours(RETR_LIST, CHAIN_APPROX_NONE) order."""
    mask = np.ascontiguousarray(mask, np.uint8)
    H, W = mask.shape
    pts = C.POINTER(C.c_int32)()
    offs = C.POINTER(C.c_int32)()
    nc = lib().orc_find_contours(_p(mask), W, H, C.byref(pts), C.byref(offs))
    o = np.ctypeslib.as_array(offs, shape=(nc + 1,)).copy()
    total = int(o[-1])
    P = np.ctypeslib.as_array(pts, shape=(max(total, 1) * 2,)).copy().reshape(-1, 2)[:total]
    lib().orc_free(pts)
    lib().orc_free(offs)
    return [P[o[i]:o[i + 1]] for i in range(nc)]


def approx_poly_dp(pts, eps):
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    out = np.empty_like(pts)
    n = lib().orc_approx_poly_dp(_p(pts), len(pts), C.c_double(eps), _p(out))
    return out[:n].copy()


def is_contour_convex(pts):
    pts = np.ascontiguousarray(pts, np.int32).reshape(-1, 2)
    return bool(lib().orc_is_contour_convex(_p(pts), len(pts)))


def point_polygon_test(poly, pt):
    poly = np.ascontiguousarray(poly, np.float32).reshape(-1, 2)
    return int(lib().orc_point_polygon_test(_p(poly), len(poly), C.c_float(pt[0]), C.c_float(pt[1])))


def get_perspective_transform(src, dst):
    src = np.ascontiguousarray(src, np.float32).reshape(4, 2)
    dst = np.ascontiguousarray(dst, np.float32).reshape(4, 2)
    H = np.empty(9, np.float64)
    lib().orc_get_perspective_transform(_p(src), _p(dst), _p(H))
    return H.reshape(3, 3)


def warp_nearest(gray, H, S):
    gray = np.ascontiguousarray(gray, np.uint8)
    Hh, Ww = gray.shape
    Hm = np.ascontiguousarray(H, np.float64).reshape(9)
    out = np.empty((S, S), np.uint8)
    lib().orc_warp_nearest(_p(gray), Ww, Hh, _p(Hm), S, _p(out))
    return out


def otsu(img):
    img = np.ascontiguousarray(img, np.uint8)
    return int(lib().orc_otsu(_p(img), img.size))


def identify_one(gray, corners, dic, params=None):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    c = np.ascontiguousarray(corners, np.float32).reshape(8)
    d = make_dict(dic)
    p = params or default_params()
    n = dic.marker_size + 2 * p.markerBorderBits
    bits = np.zeros((n, n), np.uint8)
    mid, rot = C.c_int(-1), C.c_int(0)
    ok = lib().orc_identify_one(_p(gray), W, H, _p(c), C.byref(d), C.byref(p), C.byref(mid), C.byref(rot), _p(bits))
    return bool(ok), mid.value, rot.value, bits


def corner_subpix(gray, corners, win, max_iter=30, eps=0.1):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    c = np.array(corners, np.float32).reshape(-1, 2).copy()
    lib().orc_corner_subpix(_p(gray), W, H, _p(c), len(c), int(win), int(max_iter), C.c_double(eps))
    return c


def refine_candidate_lines(contour, corners):
    """CORNER_REFINE_CONTOUR of one marker: contour (n, 2) int, corners (4, 2) -> refined (4, 2) f32 (None when cv2 would raise)"""
    ct = np.ascontiguousarray(contour, np.int32).reshape(-1, 2)
    c = np.array(corners, np.float32).reshape(4, 2).copy()
    ok = lib().orc_refine_candidate_lines(_p(ct), len(ct), _p(c))
    return c if ok else None


def detect(img, dic, params=None, debug=False):
    """-> (corners (n,4,2) f32, ids (n,) i32, rejected (m,4,2) f32[, debug dict])"""
    img = np.ascontiguousarray(img, np.uint8)
    ch = 1 if img.ndim == 2 else img.shape[2]
    H, W = img.shape[:2]
    d = make_dict(dic)
    p = params or default_params()
    out = Detections()
    lib().orc_detect(_p(img), W, H, ch, C.byref(d), C.byref(p), C.byref(out))

    def arr(ptr, n, shape, dt):
        if n == 0:
            return np.zeros((0,) + shape, dt)
        return np.ctypeslib.as_array(ptr, shape=(n,) + shape).copy()

    corners = arr(out.corners, out.n_acc, (4, 2), np.float32)
    ids = arr(out.ids, out.n_acc, (), np.int32)
    rej = arr(out.rejected, out.n_rej, (4, 2), np.float32)
    dbg = None
    if debug:
        dbg = dict(cand=arr(out.cand, out.n_cand, (4, 2), np.float32),
                   cand_len=arr(out.cand_len, out.n_cand, (), np.int32),
                   sel=arr(out.sel, out.n_sel, (4, 2), np.float32),
                   sel_info=arr(out.sel_info, out.n_sel, (5,), np.int32),
                   n_contours=arr(out.n_contours, out.n_scales, (), np.int32))
    lib().orc_free_detections(C.byref(out))
    return (corners, ids, rej, dbg) if debug else (corners, ids, rej)


def estimate_pose_single_markers(corners, marker_length, K, D):
    c = np.ascontiguousarray(corners, np.float32).reshape(-1, 8)
    n = len(c)
    K = np.ascontiguousarray(K, np.float64).reshape(9)
    D = np.ascontiguousarray(D, np.float64).reshape(-1)
    r = np.zeros((n, 3))
    t = np.zeros((n, 3))
    lib().orc_estimate_pose_single_markers(_p(c), n, C.c_double(marker_length), _p(K), _p(D), len(D), _p(r), _p(t))
    return r, t


def rodrigues(rvec):
    r = np.ascontiguousarray(rvec, np.float64).reshape(3)
    R = np.empty(9)
    lib().orc_rodrigues(_p(r), _p(R))
    return R.reshape(3, 3)


def project_points(obj, rvec, tvec, K, D):
    obj = np.ascontiguousarray(obj, np.float64).reshape(-1, 3)
    D5 = np.zeros(5)
    D = np.asarray(D, np.float64).reshape(-1)
    D5[:min(5, len(D))] = D[:5]
    out = np.empty((len(obj), 2))
    lib().orc_project_points(_p(obj), len(obj), _p(np.ascontiguousarray(rvec, np.float64).reshape(3)),
                             _p(np.ascontiguousarray(tvec, np.float64).reshape(3)),
                             _p(np.ascontiguousarray(K, np.float64).reshape(9)), _p(D5), _p(out))
    return out


def slam_params(**kw) -> SlamParams:
    """defaults: reference parameters.yaml:5-17 and aruco_slam.h:58"""
    d = dict(Q_k=0.01, R_x=100.0, R_y=100.0, R_theta=10.0, kl=0.05, kr=0.05, b=0.09,
             marker_length=0.27, r2c_tx=0.0, r2c_ty=0.0, useful_distance_threshold=3.0)
    d.update(kw)
    return SlamParams(**d)


def make_observations(corners, ids, rvecs, tvecs, K, D, sp: SlamParams):
    c = np.ascontiguousarray(corners, np.float32).reshape(-1, 8)
    n = len(c)
    ids = np.ascontiguousarray(ids, np.int32).reshape(-1)
    D5 = np.zeros(5)
    D = np.asarray(D, np.float64).reshape(-1)
    D5[:min(5, len(D))] = D[:5]
    out = (Observation * max(n, 1))()
    k = lib().orc_make_observations(_p(c), _p(ids), _p(np.ascontiguousarray(rvecs, np.float64)),
                                    _p(np.ascontiguousarray(tvecs, np.float64)), n,
                                    _p(np.ascontiguousarray(K, np.float64).reshape(9)), _p(D5), C.byref(sp), out)
    return [out[i] for i in range(k)]


class Ekf:
    def __init__(self, sp: SlamParams):
        self._h = lib().orc_ekf_create(C.byref(sp))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ekf_destroy(self._h)
            self._h = None

    @property
    def dim(self):
        return lib().orc_ekf_dim(self._h)

    def get_state(self):
        N = self.dim
        mu = np.empty(N)
        sg = np.empty((N, N))
        ids = np.empty(max((N - 3) // 3, 1), np.int32)
        lib().orc_ekf_get_state.argtypes = [C.c_void_p] * 4
        lib().orc_ekf_get_state(self._h, _p(mu), _p(sg), _p(ids))
        return mu, sg, ids[:(N - 3) // 3]

    def set_state(self, mu, sigma, ids):
        mu = np.ascontiguousarray(mu, np.float64)
        sigma = np.ascontiguousarray(sigma, np.float64)
        ids = np.ascontiguousarray(ids, np.int32)
        lib().orc_ekf_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib().orc_ekf_set_state(self._h, len(mu), _p(mu), _p(sigma), _p(ids) if len(ids) else None)

    def predict(self, wl, wr, dt):
        lib().orc_ekf_predict.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        lib().orc_ekf_predict(self._h, wl, wr, dt)

    def update(self, observations, dense=False):
        n = len(observations)
        arr = (Observation * max(n, 1))(*observations)
        lib().orc_ekf_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib().orc_ekf_update(self._h, arr, n, int(dense))
