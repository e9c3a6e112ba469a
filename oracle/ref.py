"""ctypes binding of oracle/_ref/libarucoslam_ref.so -- the reference's own src/aruco_slam.cpp and
src/map_loader.cpp compiled unmodified against stand-in headers (oracle/ref_backend.h, oracle/Makefile).

TEST INFRASTRUCTURE ONLY: tests/, tools/make_golden_slam.py and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.  The library is built in the authoring
container (where /root/reference exists) and travels to the GPU box as a prebuilt file.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libarucoslam_ref.so")
REFERENCE = os.environ.get("B2A_REFERENCE", "/root/reference")


class Init(C.Structure):
    _fields_ = [("Q_k", C.c_double), ("R_x", C.c_double), ("R_y", C.c_double), ("R_theta", C.c_double),
                ("kl", C.c_double), ("kr", C.c_double), ("b", C.c_double), ("marker_length", C.c_double),
                ("r2c_t", C.c_double * 3), ("r2c_q", C.c_double * 4),
                ("markers_dictionary", C.c_int), ("useful_distance_threshold", C.c_float)]


DETECT_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int)
POSE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                      C.POINTER(C.c_double), C.POINTER(C.c_double))
RODRIGUES_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double))
PROJECT_FN = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                         C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_float))

_lib = None
_hooks = None          # keeps the CFUNCTYPE objects alive


def available() -> bool:
    return os.path.exists(SO) or os.path.exists(os.path.join(REFERENCE, "src", "aruco_slam.cpp"))


def build() -> str:
    if os.path.exists(os.path.join(REFERENCE, "src", "aruco_slam.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REF=" + REFERENCE])
    if not os.path.exists(SO):
        raise RuntimeError("oracle/_ref/libarucoslam_ref.so is missing and %s is not present to build it" % REFERENCE)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ref_create.restype = C.c_void_p
        for name in ("ref_destroy", "ref_dim", "ref_is_init"):
            getattr(_lib, name).argtypes = [C.c_void_p]
        _lib.ref_set_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.ref_add_encoder.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _lib.ref_add_image.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.ref_get_observations.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4
        _lib.ref_get_state.argtypes = [C.c_void_p] * 3
        _lib.ref_get_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.ref_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.ref_get_last_observed.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.ref_robot_pose.argtypes = [C.c_void_p] * 4
        _lib.ref_markers.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        _lib.ref_map_load.argtypes = [C.c_char_p, C.c_int] + [C.c_void_p] * 5
        _lib.ref_set_clock.argtypes = [C.c_double]
        _lib.ref_set_replay.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.ref_set_dictionary.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.ref_set_hooks.argtypes = [C.c_void_p] * 4
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def set_dictionary(dic):
    """dictionary for the default (oracle/orc_detect.c) detector hook; dic: aruco_slam_b200.dictionaries.Dictionary"""
    t = np.ascontiguousarray(dic.table, np.uint8)
    lib().ref_set_dictionary(dic.marker_size, dic.max_correction_bits, dic.n_markers, dic.n_bytes, _p(t))


def use_orc_hooks():
    global _hooks
    lib().ref_set_hooks(None, None, None, None)
    _hooks = None


def use_cv2_hooks():
    """route the five OpenCV calls of aruco_slam.cpp into the cv2 wheel (authoring container / any box with cv2)."""
    global _hooks
    import cv2
    dets = {}

    def detect(img, w, h, ch, dict_id, corners, ids, cap):
        a = np.ctypeslib.as_array(img, (h, w, ch) if ch == 3 else (h, w))
        if dict_id not in dets:
            dets[dict_id] = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(dict_id), cv2.aruco.DetectorParameters())
        c, i, _ = dets[dict_id].detectMarkers(a)
        n = 0 if i is None else len(i)
        for k in range(min(n, cap)):
            ids[k] = int(i[k][0])
            q = c[k].reshape(8)
            for j in range(8):
                corners[8 * k + j] = float(q[j])
        return min(n, cap)

    def pose(corners, n, L, K9, D, nD, rvecs, tvecs):
        K = np.ctypeslib.as_array(K9, (9,)).reshape(3, 3).copy()
        Dv = np.ctypeslib.as_array(D, (5,))[:nD].copy()
        h = np.float32(L) / np.float32(2)
        obj = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float32)      # estimatePoseSingleMarkers' object points
        c = np.ctypeslib.as_array(corners, (n, 4, 2))
        for k in range(n):
            ok, r, t = cv2.solvePnP(obj, c[k].reshape(-1, 1, 2).astype(np.float32), K, Dv)
            for j in range(3):
                rvecs[3 * k + j] = float(r[j, 0])
                tvecs[3 * k + j] = float(t[j, 0])

    def rodrigues(rvec, R9):
        R, _ = cv2.Rodrigues(np.ctypeslib.as_array(rvec, (3,)).copy())
        for j in range(9):
            R9[j] = float(R.reshape(9)[j])

    def project(obj, n, rvec, tvec, K9, D, nD, out):
        o = np.ctypeslib.as_array(obj, (n, 3)).astype(np.float32)
        K = np.ctypeslib.as_array(K9, (9,)).reshape(3, 3).copy()
        Dv = np.ctypeslib.as_array(D, (5,))[:nD].copy()
        p, _ = cv2.projectPoints(o, np.ctypeslib.as_array(rvec, (3,)).copy(), np.ctypeslib.as_array(tvec, (3,)).copy(), K, Dv)
        p = p.reshape(-1).astype(np.float32)
        for j in range(2 * n):
            out[j] = float(p[j])

    _hooks = (DETECT_FN(detect), POSE_FN(pose), RODRIGUES_FN(rodrigues), PROJECT_FN(project))
    lib().ref_set_hooks(*[C.cast(h, C.c_void_p) for h in _hooks])


def set_clock(t: float):
    lib().ref_set_clock(float(t))


def set_replay(corners=None, ids=None, rvecs=None, tvecs=None):
    """next detectMarkers / estimatePoseSingleMarkers calls return these arrays; no arguments = replay off"""
    if ids is None:
        z = np.zeros(1)
        lib().ref_set_replay(_p(z), _p(z), -1, _p(z), _p(z))
        return
    c = np.ascontiguousarray(corners, np.float32).reshape(-1, 8)
    i = np.ascontiguousarray(ids, np.int32).reshape(-1)
    r = np.ascontiguousarray(rvecs, np.float64).reshape(-1, 3)
    t = np.ascontiguousarray(tvecs, np.float64).reshape(-1, 3)
    lib().ref_set_replay(_p(c), _p(i), len(i), _p(r), _p(t))


class RefSlam:
    """ArucoSlam of the reference (src/aruco_slam.cpp), driven through the harness."""

    def __init__(self, Q_k=0.01, R_x=100.0, R_y=100.0, R_theta=10.0, kl=0.05, kr=0.05, b=0.09, marker_length=0.27,
                 r2c_t=(0.0, 0.0, 0.0), r2c_q=(0.0, 0.0, 0.0, 1.0), markers_dictionary=16, useful_distance_threshold=3.0):
        ini = Init(Q_k, R_x, R_y, R_theta, kl, kr, b, marker_length, (C.c_double * 3)(*r2c_t), (C.c_double * 4)(*r2c_q),
                   markers_dictionary, useful_distance_threshold)
        self._h = lib().ref_create(C.byref(ini))

    def close(self):
        if getattr(self, "_h", None):
            lib().ref_destroy(self._h)
            self._h = None

    __del__ = close

    def set_camera(self, K, D):
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        D = np.ascontiguousarray(D, np.float64).reshape(-1)
        lib().ref_set_camera(self._h, _p(K), _p(D), len(D))

    def add_encoder(self, wl, wr, t=None):
        if t is not None:
            set_clock(t)
        lib().ref_add_encoder(self._h, float(wl), float(wr))

    @staticmethod
    def _img(img):
        img = np.ascontiguousarray(img, np.uint8)
        return img, img.shape[1], img.shape[0], (1 if img.ndim == 2 else img.shape[2])

    def add_image(self, img):
        a, w, h, ch = self._img(img)
        lib().ref_add_image(self._h, _p(a), w, h, ch)

    def get_observations(self, img, cap=1024):
        a, w, h, ch = self._img(img)
        ids, idx = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        xyt, cov = np.zeros((cap, 3)), np.zeros((cap, 9))
        n = lib().ref_get_observations(self._h, _p(a), w, h, ch, cap, _p(ids), _p(idx), _p(xyt), _p(cov))
        return ids[:n].copy(), idx[:n].copy(), xyt[:n].copy(), cov[:n].copy()

    @property
    def dim(self):
        return lib().ref_dim(self._h)

    @property
    def is_init(self):
        return bool(lib().ref_is_init(self._h))

    def get_state(self):
        N = self.dim
        mu, sg = np.empty(N), np.empty((N, N))
        ids = np.full(max((N - 3) // 3, 1), -1, np.int32)
        lib().ref_get_state(self._h, _p(mu), _p(sg))
        n = lib().ref_get_ids(self._h, _p(ids), len(ids))
        return mu, sg, ids[:n]

    def set_state(self, mu, sigma, ids, is_init=True):
        mu = np.ascontiguousarray(mu, np.float64)
        sigma = np.ascontiguousarray(sigma, np.float64)
        ids = np.ascontiguousarray(ids, np.int32)
        lib().ref_set_state(self._h, len(mu), _p(mu), _p(sigma), _p(ids), int(is_init))

    def last_observed(self, cap=1024):
        ids, lo = np.zeros(cap, np.int32), np.zeros((cap, 3))
        n = lib().ref_get_last_observed(self._h, cap, _p(ids), _p(lo))
        return ids[:n].copy(), lo[:n].copy()

    def robot_pose(self):
        p, q, c = np.zeros(3), np.zeros(4), np.zeros(36)
        lib().ref_robot_pose(self._h, _p(p), _p(q), _p(c))
        return p, q, c

    def markers(self, which, cap=2048):
        i, s, p, q, c, lt = np.zeros(cap, np.int32), np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros((cap, 4)), np.zeros((cap, 4)), np.zeros(cap)
        n = lib().ref_markers(self._h, which, cap, _p(i), _p(s), _p(p), _p(q), _p(c), _p(lt))
        return dict(id=i[:n].copy(), scale=s[:n].copy(), position=p[:n].copy(), orientation=q[:n].copy(), color=c[:n].copy(), lifetime=lt[:n].copy())


def map_load(path, cap=4096):
    """MapLoader(path).toRosRealMapMarkers() of the reference (map_loader.cpp)."""
    i, s, p, q, c = np.zeros(cap, np.int32), np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros((cap, 4)), np.zeros((cap, 4))
    n = lib().ref_map_load(os.fsencode(path), cap, _p(i), _p(s), _p(p), _p(q), _p(c))
    return dict(id=i[:n].copy(), scale=s[:n].copy(), position=p[:n].copy(), orientation=q[:n].copy(), color=c[:n].copy())
