#!/usr/bin/env python
"""Write tests/golden/contour_refine.npz from the cv2 4.13.0 wheel (authoring container only): what
`ArucoDetector.detectMarkers` returns with `cornerRefinementMethod = CORNER_REFINE_CONTOUR` on the frames the
`detect_*.npz` fixtures already hold (keys `<fixture>/corners`, `<fixture>/ids`, `<fixture>/rejected`).

cv2 fits the side lines in float32 normal equations and computes A^T B with its BLAS once a side has 100 points or
more, so its corners carry up to a few hundredths of a pixel of build-dependent rounding there; the tests ask for exact
equality where every side of a marker is shorter than 90 px and for 0.05 px elsewhere.

tests/golden/refine_board.npz: `ArucoDetector.refineDetectedMarkers` on a rendered GridBoard (5 x 4 markers of DICT_6X6_250 seen
through a distorting pinhole camera, five markers damaged cell by cell until detectMarkers rejects them): the detectMarkers
output it starts from and, for several RefineParameters with and without camera, the corners / ids / rejected / recoveredIdxs
cv2 returns (keys `<case>/...`).

Run:  python tools/make_golden_refine.py        (needs cv2)
"""
import glob
import os
import sys

import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
A = cv2.aruco
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
PROV = "cv2 %s (opencv-python-headless), tools/make_golden_refine.py" % cv2.__version__


BOARD_K = np.array([[900.0, 0, 480], [0, 900.0, 270], [0, 0, 1]])
BOARD_D = np.array([-0.12, 0.05, 0.001, -0.0015, 0.0])
# (name, minRepDistance, errorCorrectionRate, checkAllOrders, with camera)
REFINE_CASES = [("h_default", 10.0, 3.0, True, False), ("cam_default", 10.0, 3.0, True, True), ("h_noorders", 10.0, 3.0, False, False),
                ("h_nocode", 10.0, -1.0, True, False), ("cam_strict", 10.0, 1.0, True, True), ("h_near", 2.0, 3.0, True, False),
                ("cam_near", 1.0, 3.0, True, True), ("h_half", 0.5, 3.0, True, False), ("cam_half", 0.45, 3.0, True, True)]


def render_board():
    """960 x 540 view of the board: every pixel's ray (undistorted with cv2) meets the board plane, the board image is sampled there"""
    dic = A.getPredefinedDictionary(A.DICT_6X6_250)
    board = A.GridBoard((5, 4), 0.04, 0.01, dic)
    bimg = board.generateImage((1000, 800), marginSize=20)             # pixel = 20 + 4000 * metre, y down
    W, H = 960, 540
    rvec = np.array([0.35, -0.42, 0.12])
    tvec = np.array([-0.13, -0.10, 0.42])
    R = cv2.Rodrigues(rvec)[0]
    uv = np.stack(np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64)), -1).reshape(-1, 1, 2)
    n = cv2.undistortPoints(uv, BOARD_K, BOARD_D).reshape(-1, 2)
    rays = np.c_[n, np.ones(len(n))]
    # board plane: X = R^T (s ray - t), z = 0
    Rt = R.T
    a = rays @ Rt.T
    b = Rt @ tvec
    s_ = b[2] / a[:, 2]
    X = a * s_[:, None] - b
    mapx = (20 + 4000 * X[:, 0]).reshape(H, W).astype(np.float32)
    mapy = (20 + 4000 * X[:, 1]).reshape(H, W).astype(np.float32)
    bg = (110 + 30 * np.sin(np.arange(W)[None, :] / 97) * np.cos(np.arange(H)[:, None] / 71)).astype(np.uint8)
    view = cv2.remap(bimg, mapx, mapy, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    inside = (mapx >= 0) & (mapx <= 999) & (mapy >= 0) & (mapy <= 799)
    frame = np.where(inside, view, bg).astype(np.uint8)
    return dic, board, frame


def board_fixture():
    dic, board, frame = render_board()
    det = A.ArucoDetector(dic, A.DetectorParameters())
    c, ids, _ = det.detectMarkers(frame)
    assert len(ids) == 20
    quads = {int(i): q.reshape(4, 2) for q, i in zip(c, ids.ravel())}
    rng = np.random.default_rng(5)
    damaged = frame.copy()
    for k, mid in enumerate([3, 7, 12, 16, 18]):                       # flip 5, 7, ... inner cells of five markers
        Hq = cv2.getPerspectiveTransform(np.array([[0, 0], [8, 0], [8, 8], [0, 8]], np.float32), quads[mid].astype(np.float32))
        for cidx in rng.permutation(36)[:5 + 2 * k]:
            i, j = divmod(int(cidx), 6)
            sq = np.array([[[j + 1, i + 1], [j + 2, i + 1], [j + 2, i + 2], [j + 1, i + 2]]], np.float32)
            pq = cv2.perspectiveTransform(sq, Hq)[0]
            ctr = pq.mean(0)
            v = 255 if frame[int(ctr[1]), int(ctr[0])] < 128 else 0
            cv2.fillConvexPoly(damaged, np.round(pq).astype(np.int32), int(v))
    c, ids, rej = det.detectMarkers(damaged)
    kw = dict(frame=damaged, dict_id=np.int32(A.DICT_6X6_250), board_ids=np.array(board.getIds(), np.int32).ravel(),
              board_obj=np.array(board.getObjPoints(), np.float32), K=BOARD_K, D=BOARD_D,
              corners=np.array(c, np.float32).reshape(-1, 4, 2), ids=ids.ravel().astype(np.int32), rejected=np.array(rej, np.float32).reshape(-1, 4, 2),
              cases=np.array([n for n, *_ in REFINE_CASES]), case_params=np.array([[a, b, float(o), float(cm)] for _, a, b, o, cm in REFINE_CASES]))
    print("refine_board: detected %d, rejected %d, missing %s" % (len(ids), len(rej), sorted(set(range(20)) - set(ids.ravel().tolist()))))
    for name, rep, ecr, orders, cam in REFINE_CASES:
        rp = A.RefineParameters(rep, ecr, orders)
        d2 = A.ArucoDetector(dic, A.DetectorParameters(), rp)
        cc, ii, rr = list(c), ids.copy(), list(rej)
        out = d2.refineDetectedMarkers(damaged, board, cc, ii, rr, cameraMatrix=BOARD_K if cam else None, distCoeffs=BOARD_D if cam else None)
        oc = np.array(out[0], np.float32).reshape(-1, 4, 2)
        oi = np.asarray(out[1]).ravel().astype(np.int32)
        orj = np.array(out[2], np.float32).reshape(-1, 4, 2)
        rec = np.zeros(0, np.int32) if out[3] is None else np.asarray(out[3]).ravel().astype(np.int32)
        kw.update({name + "/corners": oc, name + "/ids": oi, name + "/rejected": orj, name + "/recovered": rec})
        print("  %-12s -> detected %d, rejected %d, recovered candidates %s as ids %s" % (name, len(oi), len(orj), rec.tolist(), oi[len(ids):].tolist()))
    # the same with CORNER_REFINE_SUBPIX in the detector: recovered candidates are refined like accepted ones
    prm = A.DetectorParameters()
    prm.cornerRefinementMethod = A.CORNER_REFINE_SUBPIX
    d3 = A.ArucoDetector(dic, prm, A.RefineParameters())
    c3, i3, r3 = d3.detectMarkers(damaged)
    out = d3.refineDetectedMarkers(damaged, board, list(c3), i3.copy(), list(r3))
    kw.update({"subpix/in_corners": np.array(c3, np.float32).reshape(-1, 4, 2), "subpix/in_ids": i3.ravel().astype(np.int32),
               "subpix/in_rejected": np.array(r3, np.float32).reshape(-1, 4, 2), "subpix/corners": np.array(out[0], np.float32).reshape(-1, 4, 2),
               "subpix/ids": np.asarray(out[1]).ravel().astype(np.int32), "subpix/rejected": np.array(out[2], np.float32).reshape(-1, 4, 2),
               "subpix/recovered": np.asarray(out[3]).ravel().astype(np.int32)})
    print("  subpix       -> detected %d, recovered %s" % (len(out[1]), np.asarray(out[3]).ravel().tolist()))
    path = os.path.join(OUT, "refine_board.npz")
    np.savez_compressed(path, provenance=np.array(PROV), **kw)
    print("refine_board.npz %.1f KB" % (os.path.getsize(path) / 1024))


def nonplanar_fixture():
    """tests/golden/refine_nonplanar.npz: cv2.solvePnP (ITERATIVE) on object points in general position (its DLT start) and on nearly
    planar ones (its homography start), and refineDetectedMarkers with a camera on a board whose markers sit 2.5 mm above / below
    the sheet in turn (general position by cv2's test, W[2] / W[1] >= 1e-3), on the frame and detections of refine_board.npz"""
    G = np.load(os.path.join(OUT, "refine_board.npz"))
    K, D = G["K"], G["D"]
    rng = np.random.default_rng(11)
    kw = {}
    n_pnp = 0
    for n, zscale in ((6, 1.0), (7, 0.3), (12, 1.0), (20, 0.1), (24, 0.04), (24, 0.02), (40, 0.5), (16, 0.0), (9, 0.005)):
        obj = rng.uniform(-0.2, 0.2, (n, 3))
        obj[:, 2] *= zscale
        obj = obj.astype(np.float32).astype(np.float64)
        rv = rng.uniform(-0.6, 0.6, 3)
        tv = np.array([rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), rng.uniform(0.6, 1.5)])
        img = (cv2.projectPoints(obj, rv, tv, K, D)[0].reshape(-1, 2) + rng.normal(0, 0.3, (n, 2))).astype(np.float32).astype(np.float64)
        ok, r, t = cv2.solvePnP(obj, img, K, D)
        assert ok
        kw.update({"pnp/%d/obj" % n_pnp: obj, "pnp/%d/img" % n_pnp: img, "pnp/%d/rvec" % n_pnp: r.ravel(), "pnp/%d/tvec" % n_pnp: t.ravel()})
        n_pnp += 1
    dic = A.getPredefinedDictionary(int(G["dict_id"]))
    obj = G["board_obj"].copy()
    obj[::2, :, 2] += 0.0025
    obj[1::2, :, 2] -= 0.0025
    board = A.Board(obj, dic, G["board_ids"].reshape(-1, 1))
    c = [q.reshape(1, 4, 2) for q in G["corners"]]
    rej = [q.reshape(1, 4, 2) for q in G["rejected"]]
    cases = [("cam_default", 10.0, 3.0, True), ("cam_near", 3.0, 3.0, True), ("cam_nocode", 10.0, -1.0, False)]
    for name, rep, ecr, orders in cases:
        d2 = A.ArucoDetector(dic, A.DetectorParameters(), A.RefineParameters(rep, ecr, orders))
        out = d2.refineDetectedMarkers(G["frame"], board, list(c), G["ids"].reshape(-1, 1).copy(), list(rej), cameraMatrix=K, distCoeffs=D)
        rec = np.zeros(0, np.int32) if out[3] is None else np.asarray(out[3]).ravel().astype(np.int32)
        kw.update({name + "/corners": np.array(out[0], np.float32).reshape(-1, 4, 2), name + "/ids": np.asarray(out[1]).ravel().astype(np.int32),
                   name + "/rejected": np.array(out[2], np.float32).reshape(-1, 4, 2), name + "/recovered": rec})
        print("  nonplanar %-12s -> detected %d, recovered %s" % (name, len(kw[name + "/ids"]), rec.tolist()))
    path = os.path.join(OUT, "refine_nonplanar.npz")
    np.savez_compressed(path, provenance=np.array(PROV), board_obj=obj.astype(np.float32), n_pnp=np.int32(n_pnp), cases=np.array([c_[0] for c_ in cases]),
                        case_params=np.array([[c_[1], c_[2], float(c_[3])] for c_ in cases]), **kw)
    print("refine_nonplanar.npz %.1f KB" % (os.path.getsize(path) / 1024))


def main():
    if "--nonplanar-only" in sys.argv:
        nonplanar_fixture()
        return
    board_fixture()
    nonplanar_fixture()
    kw = {}
    names = []
    for path in sorted(glob.glob(os.path.join(OUT, "detect_*.npz"))):
        name = os.path.basename(path)[:-4]
        g = np.load(path)
        prm = A.DetectorParameters()
        prm.cornerRefinementMethod = A.CORNER_REFINE_CONTOUR
        c, ids, rej = A.ArucoDetector(A.getPredefinedDictionary(int(g["dict_id"])), prm).detectMarkers(g["frame"])
        kw[name + "/corners"] = np.array(c, np.float32).reshape(-1, 4, 2)
        kw[name + "/ids"] = np.zeros(0, np.int32) if ids is None else ids.ravel().astype(np.int32)
        kw[name + "/rejected"] = np.array(rej, np.float32).reshape(-1, 4, 2)
        names.append(name)
        print("%-32s %3d markers" % (name, len(kw[name + "/ids"])))
    path = os.path.join(OUT, "contour_refine.npz")
    np.savez_compressed(path, provenance=np.array(PROV), fixtures=np.array(names), **kw)
    print("contour_refine.npz %.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
