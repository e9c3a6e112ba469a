"""CORNER_REFINE_CONTOUR (cv::aruco _refineCandidateLines, inside detectMarkers of reference src/aruco_slam.cpp:313 when the
parameter asks for it): the oracle against cv2 4.13 goldens (tests/golden/contour_refine.npz, tools/make_golden_refine.py)."""
import os

import numpy as np
import pytest

from aruco_slam_b200 import dictionaries as D
from oracle import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = np.load(os.path.join(GOLD, "contour_refine.npz"))
FIXTURES = [str(n) for n in REF["fixtures"]]
# cv2 sums A^T B of a side with >= 100 points in its BLAS (float, order depends on the build): exact below, 0.05 px above
EXACT_SIDE, TOL = 90.0, 0.05


def short_sides(unrefined):
    d = np.abs(unrefined - np.roll(unrefined, -1, axis=1)).max(axis=2)          # Chebyshev length of every side
    return d.max(axis=1) < EXACT_SIDE


def check_against_golden(name, corners, ids, rejected):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    assert np.array_equal(ids, REF[name + "/ids"])
    assert np.array_equal(rejected, REF[name + "/rejected"])                     # rejected candidates are not refined
    want = REF[name + "/corners"]
    assert corners.shape == want.shape
    if len(want) == 0:
        return 0, 0
    exact = short_sides(g["corners"])
    assert np.array_equal(corners[exact], want[exact])
    assert np.abs(corners - want).max() <= TOL
    assert np.abs(want - g["corners"]).max() > 0.01                               # the refinement moved something
    return int(exact.sum()), len(want)


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_contour_refine_vs_cv2(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = oracle.default_params()
    p.cornerRefinementMethod = 2
    c, ids, rej = oracle.detect(g["frame"], D.getPredefinedDictionary(int(g["dict_id"])), p)
    check_against_golden(name, c, ids, rej)


def test_exact_cases_exist():
    n_exact = sum(int(short_sides(np.load(os.path.join(GOLD, n + ".npz"))["corners"]).sum()) for n in FIXTURES if len(REF[n + "/ids"]))
    assert n_exact >= 30


def test_refine_candidate_lines_square():
    # an axis-parallel square contour: the side lines are exact, the corners stay where they are
    pts = [(x, 10) for x in range(10, 50)] + [(50, y) for y in range(10, 50)] + [(x, 50) for x in range(50, 10, -1)] + [(10, y) for y in range(50, 10, -1)]
    c = oracle.refine_candidate_lines(np.array(pts), np.array([[10, 10], [50, 10], [50, 50], [10, 50]], np.float32))
    assert np.allclose(c, [[10, 10], [50, 10], [50, 50], [10, 50]], atol=1e-3)
    # a corner that is not a contour point: cv2 raises, the restatement reports it
    assert oracle.refine_candidate_lines(np.array(pts), np.array([[10, 10], [50, 10], [50, 50], [11, 51]], np.float32)) is None


def _contours_of(frame):
    """every kept contour of the three default masks through the oracle, keyed by the point set of its approximated quad"""
    H, W = frame.shape
    lo, hi = int(0.03 * max(W, H)), int(4.0 * max(W, H))
    out = {}
    for k in (3, 13, 23):
        for c in oracle.find_contours(oracle.adaptive_threshold(frame, k, 7)):
            if lo <= len(c) <= hi:
                q = oracle.approx_poly_dp(c, len(c) * 0.03)
                if len(q) == 4:
                    out.setdefault(frozenset(map(tuple, np.asarray(q).reshape(-1, 2))), np.asarray(c).reshape(-1, 2))
    return out


@pytest.mark.parametrize("name", ["detect_vga_4x4_s2", "detect_540p_6x6_noisy_s0", "detect_720p_orig_noisy_s1", "detect_1080p_6x6_s1"])
def test_product_header_refine_lines_equals_oracle(name):
    """refine_core.h (the product's arithmetic, one lane on the host) == the oracle's restatement, bit for bit, and cv2 within its tolerance"""
    from hostemu import emu
    g = np.load(os.path.join(GOLD, name + ".npz"))
    contours = _contours_of(g["frame"])
    got = []
    for q in g["corners"]:
        c = contours[frozenset(map(tuple, q.astype(int)))]
        ok, r = emu.refine_lines(c, q)
        assert ok
        assert np.array_equal(r, oracle.refine_candidate_lines(c, q))
        got.append(r)
    check_against_golden(name, np.array(got, np.float32), REF[name + "/ids"], REF[name + "/rejected"])


def test_product_header_refine_lines_rejects():
    from hostemu import emu
    pts = [(x, 10) for x in range(10, 50)] + [(50, y) for y in range(10, 50)] + [(x, 50) for x in range(50, 10, -1)] + [(10, y) for y in range(50, 10, -1)]
    q = np.array([[10, 10], [50, 10], [50, 50], [11, 51]], np.float32)
    ok, r = emu.refine_lines(np.array(pts), q)
    assert not ok and np.array_equal(r, q)                         # corners pass through untouched
    # the contour in the other direction gives the same corners
    q = np.array([[10, 10], [50, 10], [50, 50], [10, 50]], np.float32)
    ok1, a = emu.refine_lines(np.array(pts), q)
    ok2, b = emu.refine_lines(np.array(pts[::-1]), q)
    assert ok1 and ok2 and np.allclose(a, b, atol=1e-3) and np.allclose(a, q, atol=1e-3)


# ---- the CUDA path (k_refine_contour) through the C ABI ----
@pytest.fixture(scope="module")
def aruco():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from aruco_slam_b200 import aruco as A
    return A


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_contour_refine(aruco, name):
    """cornerRefinementMethod = CORNER_REFINE_CONTOUR on the GPU: bit-exact against the oracle, and against cv2 within what cv2's
    own float sums allow; pose follows the refined corners"""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    frame = g["frame"]
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    det = aruco.ArucoDetector(dic, aruco.DetectorParameters(cornerRefinementMethod=2), max_shape=frame.shape, max_batch=2)
    r = det.detect_batch(np.stack([frame, frame]))
    p = oracle.default_params()
    p.cornerRefinementMethod = 2
    oc, oi, orj = oracle.detect(frame, dic, p)
    for b in range(2):
        assert np.array_equal(r.ids[b], oi) and np.array_equal(r.rejected[b], orj)
        assert np.array_equal(r.corners[b], oc)
        check_against_golden(name, r.corners[b], r.ids[b], r.rejected[b])
    det.close()


@pytest.mark.gpu
def test_gpu_contour_refine_pose_and_unsupported(aruco):
    from aruco_slam_b200 import synth
    g = np.load(os.path.join(GOLD, "detect_1080p_6x6_s0.npz"))
    frame = g["frame"]
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]])
    Dc = np.zeros(5)
    det = aruco.ArucoDetector(dic, aruco.DetectorParameters(cornerRefinementMethod=2), max_shape=frame.shape, max_batch=1)
    r = det.detect_pose_batch(frame, 0.1, K, Dc)
    rv, tv = oracle.estimate_pose_single_markers(r.corners[0], 0.1, K, Dc)
    assert np.abs(r.tvecs[0] - tv).max() < 1e-4 and max(synth.rvec_distance(a, b) for a, b in zip(r.rvecs[0], rv)) < 1e-4
    det.close()
    with pytest.raises(Exception) as e:
        aruco.ArucoDetector(dic, aruco.DetectorParameters(cornerRefinementMethod=3), max_shape=frame.shape, max_batch=1)
    assert "APRILTAG" in str(e.value)
