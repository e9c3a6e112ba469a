"""ctypes binding of tests/hostemu/libhostemu.so (TEST INFRASTRUCTURE ONLY): the product's
host/device kernel logic headers compiled for the host, one lane, to check them on CPU."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostemu.so")
_CSRC = os.path.join(_HERE, "..", "..", "aruco_slam_b200", "csrc")


class EmuParams(C.Structure):
    _fields_ = [("nScales", C.c_int), ("radius", C.c_int * 8), ("Cfloor", C.c_int),
                ("minPerimRate", C.c_double), ("maxPerimRate", C.c_double), ("approxRate", C.c_double),
                ("minCornerDistRate", C.c_double), ("minDistanceToBorder", C.c_int),
                ("minMarkerDistanceRate", C.c_float), ("minGroupDistance", C.c_float),
                ("markerSize", C.c_int), ("borderBits", C.c_int), ("cellSize", C.c_int), ("cellMargin", C.c_int),
                ("nMarkers", C.c_int), ("maxCorr", C.c_int), ("maxBorderErr", C.c_int), ("minOtsuStdDev", C.c_double),
                ("max_cand", C.c_int), ("max_markers", C.c_int), ("surv_cap", C.c_int), ("detectInverted", C.c_int)]


def build():
    srcs = [os.path.join(_HERE, "hostemu.cpp")] + [os.path.join(_CSRC, f) for f in ("core.h", "frame_logic.h", "pose_core.h", "draw_core.h", "refine_core.h", "board_core.h", "pyr_core.h")] + [os.path.join(_CSRC, "..", "data", "overlay_tables.inc")]
    if not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, srcs[0]])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def pack_dict(dic):
    t = dic.table.astype(np.uint64)
    out = np.zeros((dic.n_markers, 4), np.uint64)
    for k in range(dic.n_bytes):
        out |= t[:, :, k] << np.uint64(8 * k)
    return np.ascontiguousarray(out)


def params_for(dic, max_cand=2048, max_markers=256, surv_cap=4096, detect_inverted=False):
    p = EmuParams()
    p.nScales = 3
    for i, r in enumerate((1, 6, 11)):
        p.radius[i] = r
    p.Cfloor = 7
    p.minPerimRate, p.maxPerimRate, p.approxRate, p.minCornerDistRate = 0.03, 4.0, 0.03, 0.05
    p.minDistanceToBorder = 3
    p.minMarkerDistanceRate, p.minGroupDistance = 0.125, 0.21
    p.markerSize, p.borderBits, p.cellSize, p.cellMargin = dic.marker_size, 1, 4, int(0.13 * 4)
    p.nMarkers = dic.n_markers
    p.maxCorr = int(dic.max_correction_bits * 0.6)
    p.maxBorderErr = int(dic.marker_size * dic.marker_size * 0.35)
    p.minOtsuStdDev = 5.0
    p.max_cand, p.max_markers, p.surv_cap = max_cand, max_markers, surv_cap
    p.detectInverted = 1 if detect_inverted else 0
    return p


def detect(gray, dic, masks=None, dbg_scale=-1, anchor_R=32, detect_inverted=False):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    p = params_for(dic, detect_inverted=detect_inverted)
    d = pack_dict(dic)
    n_acc, n_rej, n_cand, nk = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    corners = np.zeros((p.max_markers, 4, 2), np.float32)
    ids = np.zeros(p.max_markers, np.int32)
    rej = np.zeros((p.max_markers, 4, 2), np.float32)
    ncont = np.zeros(3, np.int32)
    cand = np.zeros((p.max_cand, 4, 2), np.float32)
    cap, pcap = 8192, 2 * W * H
    dlen = np.zeros(cap, np.int32)
    dpts = np.zeros((pcap, 2), np.int16)
    mp = None
    if masks is not None:
        mk = np.ascontiguousarray(masks, np.uint8)
        mp = mk.ctypes.data_as(C.c_void_p)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    st = lib().emu_detect(P(gray), W, H, mp, P(d), C.byref(p), C.byref(n_acc), C.byref(n_rej), P(corners), P(ids), P(rej),
                          P(ncont), C.byref(n_cand), P(cand), dbg_scale, C.byref(nk), P(dlen), cap, P(dpts), pcap, anchor_R)
    out = dict(status=st, corners=corners[:n_acc.value].copy(), ids=ids[:n_acc.value].copy(), rejected=rej[:n_rej.value].copy(),
               n_contours=ncont, cand=cand[:n_cand.value].copy())
    if dbg_scale >= 0:
        ln = dlen[:nk.value].copy()
        out["kept_len"] = ln
        out["kept_pts"] = dpts[:int(ln.sum())].copy()
    return out


def pyr_down(gray):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    out = np.zeros(((H + 1) // 2, (W + 1) // 2), np.uint8)
    lib().emu_pyr_down(gray.ctypes.data_as(C.c_void_p), W, H, out.ctypes.data_as(C.c_void_p))
    return out


def resize_linear(gray, dW, dH):
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    out = np.zeros((dH, dW), np.uint8)
    lib().emu_resize_linear(gray.ctypes.data_as(C.c_void_p), W, H, out.ctypes.data_as(C.c_void_p), int(dW), int(dH))
    return out


def detect_aruco3(gray, dic, min_side=32, ratio=0.0, detect_inverted=False):
    """ArUco3 through the product's headers: ids / rejected as the device returns them, accepted corners before the refinement
    chain (segmentation-image coordinates), plan = (segW, segH, numLevels, closestIdx); None when the plan is refused"""
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    p = params_for(dic, detect_inverted=detect_inverted)
    d = pack_dict(dic)
    n_acc, n_rej = C.c_int(), C.c_int()
    corners = np.zeros((p.max_markers, 4, 2), np.float32)
    ids = np.zeros(p.max_markers, np.int32)
    rej = np.zeros((p.max_markers, 4, 2), np.float32)
    plan = np.zeros(4, np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    st = lib().emu_detect_aruco3(P(gray), W, H, P(d), C.byref(p), int(min_side), C.c_float(ratio), C.byref(n_acc), C.byref(n_rej), P(corners), P(ids), P(rej), P(plan))
    if st < 0:
        return None
    return dict(status=st, corners=corners[:n_acc.value].copy(), ids=ids[:n_acc.value].copy(), rejected=rej[:n_rej.value].copy(), plan=tuple(int(v) for v in plan))


def pose(corners, K, D, L):
    c = np.ascontiguousarray(corners, np.float32).reshape(-1, 8)
    D5 = np.zeros(5); D = np.asarray(D, float).ravel(); D5[:len(D)] = D
    r = np.zeros((len(c), 3)); t = np.zeros((len(c), 3))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().emu_pose(P(c), len(c), P(np.ascontiguousarray(K, np.float64).reshape(9)), P(D5), C.c_float(L), P(r), P(t))
    return r, t


def observation(corners, mid, rvec, tvec, K, D, sp):
    D5 = np.zeros(5); D = np.asarray(D, float).ravel(); D5[:len(D)] = D
    out = np.zeros(12)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().emu_observation.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_double] * 6 + [C.c_float, C.c_void_p]
    k = lib().emu_observation(P(np.ascontiguousarray(corners, np.float32).reshape(8)), int(mid), P(np.ascontiguousarray(rvec, np.float64)),
                              P(np.ascontiguousarray(tvec, np.float64)), P(np.ascontiguousarray(K, np.float64).reshape(9)), P(D5),
                              sp.R_x, sp.R_y, sp.R_theta, sp.marker_length, sp.r2c_tx, sp.r2c_ty, sp.useful_distance_threshold, P(out))
    return bool(k), out


def otsu(hist):
    """(threshold of the product's restated Otsu, threshold of the textbook sequential loop)"""
    h = np.ascontiguousarray(hist, np.int32)
    assert h.shape == (256,)
    a, b = C.c_int(), C.c_int()
    lib().emu_otsu(h.ctypes.data_as(C.c_void_p), int(h.sum()), C.byref(a), C.byref(b))
    return a.value, b.value


def approx(contour, eps):
    """product approxPolyDP(closed=True) on an (n,2) integer contour -> (m,2) vertices, or None when it gave up (> 8 vertices)"""
    c = np.ascontiguousarray(contour, np.int32).reshape(-1, 2)
    out = np.zeros((8, 2), np.int32)
    lib().emu_approx.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    m = lib().emu_approx(c.ctypes.data_as(C.c_void_p), len(c), float(eps), out.ctypes.data_as(C.c_void_p))
    return None if m < 0 else out[:m].copy()


def draw(image, corners, ids=None, border=(0, 255, 0)):
    """product drawDetectedMarkers logic (draw_core.h) on a copy of `image` ((H,W) or (H,W,3) uint8)"""
    img = np.ascontiguousarray(image, np.uint8).copy()
    H, W = img.shape[:2]
    ch = 1 if img.ndim == 2 else 3
    c = np.ascontiguousarray(corners, np.float32).reshape(-1, 8)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    idp = None
    if ids is not None:
        ida = np.ascontiguousarray(ids, np.int32).reshape(-1)
        idp = P(ida)
    b = np.array(border, np.uint8)
    lib().emu_draw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib().emu_draw(P(img), W, H, ch, P(c), idp, len(c), P(b))
    return img


def refine_lines(contour, corners):
    """CORNER_REFINE_CONTOUR of one marker through refine_core.h: (ok, refined (4, 2) f32)"""
    ct = np.ascontiguousarray(contour, np.int32).reshape(-1, 2)
    cin = np.ascontiguousarray(corners, np.float32).reshape(8)
    out = np.zeros(8, np.float32)
    ok = lib().emu_refine_lines(ct.ctypes.data_as(C.c_void_p), len(ct), cin.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return bool(ok), out.reshape(4, 2)


def homography(src, dst, refine_iters=10):
    s = np.ascontiguousarray(src, np.float64).reshape(-1, 2)
    d = np.ascontiguousarray(dst, np.float64).reshape(-1, 2)
    H = np.zeros(9)
    ok = lib().emu_homography(s.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), len(s), H.ctypes.data_as(C.c_void_p), int(refine_iters))
    return bool(ok), H.reshape(3, 3)


def board_pose(K, D, obj, img):
    K9 = np.ascontiguousarray(K, np.float64).reshape(9)
    D5 = np.r_[np.asarray(D, np.float64).ravel(), np.zeros(5)][:5].copy()
    o = np.ascontiguousarray(obj, np.float64).reshape(-1, 3)
    i = np.ascontiguousarray(img, np.float64).reshape(-1, 2)
    r, t = np.zeros(3), np.zeros(3)
    rc = lib().emu_board_pose(K9.ctypes.data_as(C.c_void_p), D5.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p), len(o),
                              r.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
    return rc, r, t
