#!/usr/bin/env python
"""Write tests/golden/*.npz from the cv2 4.13.0 wheel (authoring container only).

The reference (gitAugust/Aruco_Slam) has no tests or golden vectors (SURVEY
section 4); its detect+pose arithmetic is OpenCV's (src/aruco_slam.cpp:313-314).
This script records what the installed `cv2` -- the runnable third-party
implementation -- returns on committed input frames, stage by stage, so that
the CPU oracle (oracle/) and through it the CUDA path are pinned to it.

Run:  python tools/make_golden.py       (needs cv2; not available/needed on the GPU box)
"""
import os
import sys
import zlib

import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from aruco_slam_b200 import synth, dictionaries as D  # noqa: E402

A = cv2.aruco
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
PROV = "cv2 %s (opencv-python-headless), tools/make_golden.py" % cv2.__version__


def cv_detect(img, dict_id, subpix=False, inverted=False):
    prm = A.DetectorParameters()
    if subpix:
        prm.cornerRefinementMethod = A.CORNER_REFINE_SUBPIX
    prm.detectInvertedMarker = bool(inverted)
    det = A.ArucoDetector(A.getPredefinedDictionary(dict_id), prm)
    c, ids, rej = det.detectMarkers(img)
    c = np.array(c, np.float32).reshape(-1, 4, 2)
    ids = np.zeros(0, np.int32) if ids is None else ids.ravel().astype(np.int32)
    rej = np.array(rej, np.float32).reshape(-1, 4, 2)
    return c, ids, rej


def cv_masks(gray):
    return [cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, k, 7)
            for k in (3, 13, 23)]


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, provenance=np.array(PROV), **kw)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def stage_fixture(name, gray, dict_id):
    """all stage outputs for one (small) frame"""
    H, W = gray.shape
    bgr = synth.gray_to_bgr(gray, 7)                  # regenerated from `frame` in the tests
    gray2 = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    kw = dict(frame=gray, bgr_crop=bgr[100:220, 200:360].copy(), bgr_gray_crop=gray2[100:220, 200:360].copy(),
              bgr_gray_crc=np.uint64(zlib.crc32(gray2.tobytes())), dict_id=dict_id)
    minp = int(0.03 * max(W, H))
    maxp = int(4.0 * max(W, H))
    for si, m in enumerate(cv_masks(gray)):
        kw["mask%d" % si] = np.packbits(m > 0)
        cs, _ = cv2.findContours(m, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        offs = np.cumsum([0] + [len(c) for c in cs]).astype(np.int32)
        pts = (np.concatenate([c.reshape(-1, 2) for c in cs]) if cs else np.zeros((0, 2))).astype(np.int16)
        kw["cont_offs%d" % si] = offs
        kw["cont_pts%d" % si] = pts
        ap_idx, ap_offs, ap_pts = [], [0], []
        for i, c in enumerate(cs):
            n = len(c)
            if n < minp or n > maxp:
                continue
            a = cv2.approxPolyDP(c, n * 0.03, True).reshape(-1, 2)
            ap_idx.append(i)
            ap_pts.append(a)
            ap_offs.append(ap_offs[-1] + len(a))
        kw["approx_idx%d" % si] = np.array(ap_idx, np.int32)
        kw["approx_offs%d" % si] = np.array(ap_offs, np.int32)
        kw["approx_pts%d" % si] = (np.concatenate(ap_pts) if ap_pts else np.zeros((0, 2))).astype(np.int16)
    c, ids, rej = cv_detect(gray, dict_id)
    kw.update(corners=c, ids=ids, rejected=rej)
    cb, idb, rejb = cv_detect(bgr, dict_id)
    kw.update(bgr_corners=cb, bgr_ids=idb, bgr_rejected=rejb)
    save(name, **kw)


def detect_fixture(name, gray, dict_id, store_mask_crc=True):
    kw = dict(frame=gray, dict_id=dict_id)
    c, ids, rej = cv_detect(gray, dict_id)
    kw.update(corners=c, ids=ids, rejected=rej)
    cs, ids2, rej2 = cv_detect(gray, dict_id, subpix=True)
    kw.update(subpix_corners=cs, subpix_ids=ids2)
    if store_mask_crc:
        ms = cv_masks(gray)
        kw["mask_crc"] = np.array([zlib.crc32(m.tobytes()) for m in ms], np.uint64)
        kw["n_contours"] = np.array([len(cv2.findContours(m, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)[0]) for m in ms], np.int32)
    save(name, **kw)


def inverted_fixture(name, gray, dict_id):
    """detectInvertedMarker = true (white markers on black): ids / corners / rejected of cv2 on the committed frame"""
    c, ids, rej = cv_detect(gray, dict_id, inverted=True)
    save(name, frame=gray, dict_id=dict_id, corners=c, ids=ids, rejected=rej)


def inverted_frames():
    """(name, frame, dictionary): a fully inverted frame, one with only its left half inverted (both kinds of marker in one
    frame), a noisy inverted 6x6 frame and the normal nested fixture under the flag (the group order is reversed by it)"""
    a = 255 - synth.render_config("C1", 6).image
    b = synth.render_config("C1", 7).image.copy()
    b[:, :320] = 255 - b[:, :320]
    cfg = dict(W=960, H=540, n_markers=12, dict_id=D.DICT_6X6_250, side_range=(50.0, 110.0), noise_sigma=4.0, blur_sigma=1.0)
    c = 255 - synth.render_frame(seed=2, **cfg).image
    return [("inverted_vga_4x4_all", a, D.DICT_4X4_50), ("inverted_vga_4x4_half", b, D.DICT_4X4_50),
            ("inverted_540p_6x6_noisy", c, D.DICT_6X6_250), ("inverted_vga_nested", nested_fixture(), D.DICT_4X4_50)]


def nested_fixture():
    """a valid small marker inside a white cell region of a big marker + a second big one
    (SURVEY probe P16) and a marker whose quiet zone touches the image border."""
    dic = D.getPredefinedDictionary(D.DICT_4X4_50)
    img = synth.background(640, 480)
    big = np.array([[60, 60], [360, 70], [350, 370], [50, 360]], float)
    synth.paste_marker(img, dic, 5, big)
    # find a white cell of marker 5 and put a small marker there: use a white quiet-zone-free patch
    small = np.array([[420, 80], [520, 85], [515, 185], [415, 180]], float)
    synth.paste_marker(img, dic, 9, small)
    inner = np.array([[150, 150], [200, 152], [198, 202], [148, 200]], float)
    synth.paste_marker(img, dic, 11, inner)
    edge = np.array([[560, 300], [632, 302], [630, 372], [558, 370]], float)
    synth.paste_marker(img, dic, 20, edge, quiet=0.05)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def pose_fixture():
    rng = np.random.default_rng(123)
    K = np.array([[1400., 0, 960], [0, 1400., 540], [0, 0, 1]])
    # reference default.yaml:16-20 style 5-term distortion (values of that order)
    Dist = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
    rows = []
    for L in (0.27, 0.1):
        obj = synth.marker_object_points(L)
        objf = obj.astype(np.float32)
        cnt = 0
        while cnt < 100:
            rv = rng.normal(0, 0.5, 3)
            rv[0] += np.pi
            tv = np.array([rng.uniform(-1, 1), rng.uniform(-.6, .6), rng.uniform(0.6, 5) * L / 0.27])
            ip = synth.project(obj, rv, tv, K, Dist)
            if (ip[:, 0] < 0).any() or (ip[:, 0] > 1919).any() or (ip[:, 1] < 0).any() or (ip[:, 1] > 1079).any():
                continue
            R = synth.rodrigues(rv)
            if R[2, 2] > -0.2:
                continue
            ipr = np.rint(ip).astype(np.float32) if cnt % 2 == 0 else ip.astype(np.float32)
            for Kd, Dd, tag in ((K, Dist, 1), (K, np.zeros(5), 0)):
                ok, r1, t1 = cv2.solvePnP(objf, ipr.reshape(-1, 1, 2), Kd, Dd)
                assert ok
                pp, _ = cv2.projectPoints(objf, r1, t1, Kd, Dd)
                Rm, _ = cv2.Rodrigues(r1)
                rows.append(np.concatenate([[L, tag], ipr.ravel(), r1.ravel(), t1.ravel(), pp.ravel(), Rm.ravel()]))
            cnt += 1
    save("pose", rows=np.array(rows), K=K, D=Dist,
         columns=np.array("L use_dist corners[8] rvec[3] tvec[3] proj[8] R[9]"))


def pose_detected_fixture():
    """poses of really detected (integer-valued, small, nearly fronto-parallel) markers: this is where
    the planar pose ambiguity bites and the Levenberg-Marquardt *schedule* decides the result."""
    K = np.array([[1400., 0, 960], [0, 1400., 540], [0, 0, 1]])
    Dist = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
    L = 0.27
    h = np.float32(L) / np.float32(2)
    obj = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float32)
    rows = []
    for seed in list(range(100, 104)) + list(range(0, 12)):
        fr = synth.render_config("C2", seed).image
        c, ids, _ = cv_detect(fr, D.DICT_6X6_250)
        for cc in c:
            ok, r1, t1 = cv2.solvePnP(obj, cc.reshape(-1, 1, 2), K, Dist)
            rows.append(np.concatenate([[L, 1], cc.ravel(), r1.ravel(), t1.ravel()]))
    save("pose_detected", rows=np.array(rows), K=K, D=Dist, columns=np.array("L use_dist corners[8] rvec[3] tvec[3]"))


def prims_fixture():
    rng = np.random.default_rng(5)
    fr = synth.render_config("C1", 11)
    img = fr.image
    quads, Hs, patches, otsus = [], [], [], []
    for t in range(60):
        j = rng.integers(len(fr.ids))
        q = np.rint(fr.corners[j] + rng.uniform(-3, 3, (4, 2))).astype(np.float32)
        S = 24
        dst = np.array([[0, 0], [S - 1, 0], [S - 1, S - 1], [0, S - 1]], np.float32)
        Hm = cv2.getPerspectiveTransform(q, dst)
        patch = cv2.warpPerspective(img, Hm, (S, S), flags=cv2.INTER_NEAREST)
        t_otsu, _ = cv2.threshold(patch, 125, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
        quads.append(q); Hs.append(Hm); patches.append(patch); otsus.append(int(t_otsu))
    cq = rng.integers(-20, 20, (4000, 4, 2)).astype(np.int32)
    convex = np.array([cv2.isContourConvex(q.reshape(-1, 1, 2)) for q in cq], np.uint8)
    pq = rng.integers(0, 30, (4000, 4, 2)).astype(np.float32)
    pp = rng.integers(0, 30, (4000, 2)).astype(np.float32)
    ppt = np.array([cv2.pointPolygonTest(q.reshape(-1, 1, 2), (float(p[0]), float(p[1])), False) for q, p in zip(pq, pp)], np.int8)
    save("prims", frame=img, quads=np.array(quads), H=np.array(Hs), patches=np.array(patches), otsu=np.array(otsus, np.int32),
         convex_quads=cq, convex=convex, ppt_quads=pq, ppt_pts=pp, ppt=ppt)


def main():
    os.makedirs(OUT, exist_ok=True)
    for s in (0, 1):
        stage_fixture("stages_vga_4x4_s%d" % s, synth.render_config("C1", s).image, D.DICT_4X4_50)
    stage_fixture("stages_vga_4x4_noisy", synth.render_frame(**{**synth.CONFIGS["C1"], "noise_sigma": 4.0, "blur_sigma": 1.0}, seed=2).image, D.DICT_4X4_50)
    for s in (2, 3, 4, 5):
        detect_fixture("detect_vga_4x4_s%d" % s, synth.render_config("C1", s).image, D.DICT_4X4_50)
    for s in (0, 1, 2):
        detect_fixture("detect_1080p_6x6_s%d" % s, synth.render_config("C2", s).image, D.DICT_6X6_250)
    cfg = dict(W=960, H=540, n_markers=12, dict_id=D.DICT_6X6_250, side_range=(50.0, 110.0), noise_sigma=4.0, blur_sigma=1.0)
    for s in (0, 1):
        detect_fixture("detect_540p_6x6_noisy_s%d" % s, synth.render_frame(seed=s, **cfg).image, D.DICT_6X6_250)
    cfg = dict(W=1280, H=720, n_markers=10, dict_id=D.DICT_ARUCO_ORIGINAL, side_range=(60.0, 140.0), noise_sigma=3.0, blur_sigma=0.8)
    for s in (0, 1):
        detect_fixture("detect_720p_orig_noisy_s%d" % s, synth.render_frame(seed=s, **cfg).image, D.DICT_ARUCO_ORIGINAL)
    detect_fixture("detect_vga_nested", nested_fixture(), D.DICT_4X4_50)
    # odd size (width not a multiple of 16) and a tiny frame
    cfg = dict(W=333, H=251, n_markers=2, dict_id=D.DICT_5X5_100, side_range=(50.0, 80.0))
    detect_fixture("detect_odd_5x5", synth.render_frame(seed=1, **cfg).image, D.DICT_5X5_100)
    detect_fixture("detect_blank", np.full((120, 160), 128, np.uint8), D.DICT_4X4_50)
    for name, img, did in inverted_frames():
        inverted_fixture(name, img, did)
    pose_fixture()
    pose_detected_fixture()
    prims_fixture()
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("# Golden vectors\n\nWritten by `tools/make_golden.py` in the authoring container from\n"
                "`%s`.\n\nThe reference repository has no tests or golden vectors of its own (SURVEY.md section 4);\n"
                "these files record what the third-party library it calls (OpenCV, `src/aruco_slam.cpp:313-314`)\n"
                "returns on the committed input frames, and pin `oracle/` to it.\n\n```\n%s\n```\n"
                % (PROV, "\n".join(l for l in cv2.getBuildInformation().split("\n")
                                   if any(k in l for k in ("Version control", "IPP", "Parallel framework", "CPU/HW", "Baseline", "Dispatched")))))


if __name__ == "__main__":
    main()
