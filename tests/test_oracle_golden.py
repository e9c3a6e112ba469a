"""Pin the CPU oracle (oracle/) to the cv2 4.13.0 outputs in tests/golden.

These are the "oracle vs golden vectors" tests: they run without a GPU and
without cv2.  The reference has no tests of its own; the
golden files are what the library it calls returns (tools/make_golden.py).
"""
import zlib

import numpy as np
import pytest

from conftest import golden, golden_names
from aruco_slam_b200 import dictionaries as D, synth


@pytest.mark.parametrize("name", golden_names("stages_"))
def test_stage_outputs(oracle, name):
    g = golden(name)
    gray = g["frame"]
    H, W = gray.shape
    # A1 gray
    bgr = synth.gray_to_bgr(gray, 7)
    assert np.array_equal(bgr[100:220, 200:360], g["bgr_crop"])
    og = oracle.bgr2gray(bgr)
    assert np.array_equal(og[100:220, 200:360], g["bgr_gray_crop"])
    assert zlib.crc32(og.tobytes()) == int(g["bgr_gray_crc"])
    minp, maxp = int(0.03 * max(W, H)), int(4.0 * max(W, H))
    for si, k in enumerate((3, 13, 23)):
        # A2 masks, bit exact
        m = oracle.adaptive_threshold(gray, k, 7.0)
        assert np.array_equal(np.packbits(m > 0), g["mask%d" % si])
        assert set(np.unique(m)) <= {0, 255}
        # A3a contours: same list, same order, same points
        cs = oracle.find_contours(m)
        offs = g["cont_offs%d" % si]
        pts = g["cont_pts%d" % si].astype(np.int32)
        assert len(cs) == len(offs) - 1
        assert np.array_equal(np.cumsum([0] + [len(c) for c in cs]), offs)
        assert np.array_equal(np.concatenate(cs) if cs else np.zeros((0, 2), np.int32), pts)
        # A3b approxPolyDP
        aoffs, apts = g["approx_offs%d" % si], g["approx_pts%d" % si].astype(np.int32)
        for j, ci in enumerate(g["approx_idx%d" % si]):
            c = cs[ci]
            assert minp <= len(c) <= maxp
            a = oracle.approx_poly_dp(c, len(c) * 0.03)
            assert np.array_equal(a, apts[aoffs[j]:aoffs[j + 1]]), (si, ci)
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    c, ids, rej = oracle.detect(gray, dic)
    assert np.array_equal(ids, g["ids"]) and np.array_equal(c, g["corners"]) and np.array_equal(rej, g["rejected"])
    c, ids, rej = oracle.detect(bgr, dic)
    assert np.array_equal(ids, g["bgr_ids"]) and np.array_equal(c, g["bgr_corners"]) and np.array_equal(rej, g["bgr_rejected"])


@pytest.mark.parametrize("name", golden_names("detect_"))
def test_detect(oracle, name):
    g = golden(name)
    gray = g["frame"]
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    c, ids, rej, dbg = oracle.detect(gray, dic, debug=True)
    assert np.array_equal(ids, g["ids"])                  # bit exact, same order
    assert np.array_equal(c, g["corners"])
    assert np.array_equal(rej, g["rejected"])
    for si, k in enumerate((3, 13, 23)):
        m = oracle.adaptive_threshold(gray, k, 7.0)
        assert zlib.crc32(m.tobytes()) == int(g["mask_crc"][si])
    assert np.array_equal(dbg["n_contours"], g["n_contours"])
    # A8 sub-pixel refinement: float32 arithmetic, tolerance 1e-3 px (north_star: 0.05 px)
    c2, ids2, _ = oracle.detect(gray, dic, oracle.default_params(cornerRefinementMethod=1))
    assert np.array_equal(ids2, g["subpix_ids"])
    if len(ids2):
        assert np.abs(c2 - g["subpix_corners"]).max() < 1e-3


@pytest.mark.parametrize("name", golden_names("inverted_"))
def test_detect_inverted(oracle, name):
    """detectInvertedMarker = true: white markers on black (and the reversed walk of every candidate group), pinned to cv2"""
    g = golden(name)
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    c, ids, rej = oracle.detect(g["frame"], dic, oracle.default_params(detectInvertedMarker=1))
    assert np.array_equal(ids, g["ids"])
    assert np.array_equal(c, g["corners"])
    assert np.array_equal(rej, g["rejected"])
    if "all" in name:                                   # without the flag none of the white markers identifies
        assert len(oracle.detect(g["frame"], dic)[1]) == 0


def test_nested_fixture_semantics(oracle):
    """SURVEY probe P16: the enclosing big marker is never identified."""
    g = golden("detect_vga_nested")
    assert sorted(g["ids"].tolist()) == [9, 11]


def test_primitives(oracle):
    g = golden("prims")
    img = g["frame"]
    S = 24
    dst = np.array([[0, 0], [S - 1, 0], [S - 1, S - 1], [0, S - 1]], np.float32)
    for q, Hm, patch, t in zip(g["quads"], g["H"], g["patches"], g["otsu"]):
        Ho = oracle.get_perspective_transform(q, dst)
        assert np.array_equal(Ho, Hm)                     # bit exact doubles
        assert np.array_equal(oracle.warp_nearest(img, Ho, S), patch)
        assert oracle.otsu(patch) == int(t)
    for q, cvx in zip(g["convex_quads"], g["convex"]):
        assert oracle.is_contour_convex(q) == bool(cvx)
    for q, p, r in zip(g["ppt_quads"], g["ppt_pts"], g["ppt"]):
        assert oracle.point_polygon_test(q, p) == int(r)


def test_pose(oracle):
    """estimatePoseSingleMarkers == per-marker solvePnP(ITERATIVE): 1e-4 rad / 1e-4 m
    (north_star tolerance); observed <= 3e-7."""
    g = golden("pose")
    K, Dd = g["K"], g["D"]
    worst_r = worst_t = 0.0
    for row in g["rows"]:
        L, use_d = row[0], int(row[1])
        corners = row[2:10].astype(np.float32)
        rvec, tvec, proj, R = row[10:13], row[13:16], row[16:24], row[24:33]
        dist = Dd if use_d else np.zeros(5)
        r, t = oracle.estimate_pose_single_markers(corners.reshape(1, 4, 2), L, K, dist)
        worst_r = max(worst_r, np.abs(r[0] - rvec).max())
        worst_t = max(worst_t, np.abs(t[0] - tvec).max())
        obj = synth.marker_object_points(float(np.float32(L)))
        obj = obj.astype(np.float32).astype(np.float64)
        assert np.abs(oracle.project_points(obj, rvec, tvec, K, dist).ravel() - proj).max() < 1e-3  # cv2 returns f32
        assert np.abs(oracle.rodrigues(rvec).ravel() - R).max() < 1e-12
    assert worst_r < 1e-4 and worst_t < 1e-4, (worst_r, worst_t)
    assert worst_r < 1e-5 and worst_t < 1e-5


def test_pose_detected_markers(oracle):
    """Detected (integer, small) quads: planar pose ambiguity; the LM schedule must pick cv2's minimum."""
    g = golden("pose_detected")
    K, Dd = g["K"], g["D"]
    rows = g["rows"]
    r, t = oracle.estimate_pose_single_markers(rows[:, 2:10].astype(np.float32).reshape(-1, 4, 2), 0.27, K, Dd)
    assert np.abs(t - rows[:, 13:16]).max() < 1e-4
    assert max(synth.rvec_distance(a, b) for a, b in zip(r, rows[:, 10:13])) < 1e-4
    assert np.abs(r - rows[:, 10:13]).max() < 5e-4            # same rotation-vector branch, too
    assert np.median(np.abs(r - rows[:, 10:13]).max(1)) < 1e-7
