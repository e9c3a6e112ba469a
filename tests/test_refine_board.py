"""refineDetectedMarkers (cv::aruco::ArucoDetector::refineDetectedMarkers, part of the detectMarkers surface of reference
src/aruco_slam.cpp:313): the numpy restatement (oracle/refine_np.py) and the product (b2a_refine_detected_markers)
against what cv2 4.13 returned on the rendered, damaged GridBoard of tests/golden/refine_board.npz (tools/make_golden_refine.py)."""
import os

import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))

from aruco_slam_b200 import dictionaries as D
from oracle import oracle, refine_np

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "refine_board.npz"))
CASES = [(str(n), tuple(p)) for n, p in zip(G["cases"], G["case_params"])]
DIC = D.getPredefinedDictionary(int(G["dict_id"]))


def same_as_cv2(name, c, i, r, rec):
    assert np.array_equal(i, G[name + "/ids"])
    assert np.array_equal(c, G[name + "/corners"])                 # recovered corners are the candidates' own, rotated: exact
    assert np.array_equal(r, G[name + "/rejected"])
    assert np.array_equal(rec, G[name + "/recovered"])


def test_golden_has_every_kind_of_case():
    n = [len(G[c + "/recovered"]) for c, _ in CASES]
    assert 0 in n and 5 in n and 3 in n and 4 in n                # nothing, everything, and two partial recoveries


def test_oracle_detect_is_the_starting_point():
    c, i, r = oracle.detect(G["frame"], DIC)
    assert np.array_equal(c, G["corners"]) and np.array_equal(i, G["ids"]) and np.array_equal(r, G["rejected"])


@pytest.mark.parametrize("name,prm", CASES)
def test_numpy_restatement_vs_cv2(name, prm):
    rep, ecr, orders, cam = prm
    out = refine_np.refine_detected_markers(G["frame"], DIC, G["board_ids"], G["board_obj"], G["corners"], G["ids"], G["rejected"],
                                            K=G["K"] if cam else None, D=G["D"] if cam else None, min_rep_distance=rep,
                                            error_correction_rate=ecr, check_all_orders=bool(orders))
    same_as_cv2(name, *out)


def test_numpy_restatement_subpix_vs_cv2():
    p = oracle.default_params()
    p.cornerRefinementMethod = 1
    c, i, r, rec = refine_np.refine_detected_markers(G["frame"], DIC, G["board_ids"], G["board_obj"], G["subpix/in_corners"], G["subpix/in_ids"],
                                                     G["subpix/in_rejected"], params=p)
    assert np.array_equal(i, G["subpix/ids"]) and np.array_equal(rec, G["subpix/recovered"]) and np.array_equal(r, G["subpix/rejected"])
    assert np.abs(c - G["subpix/corners"]).max() < 0.05 and len(rec) == 5
    n0 = len(G["subpix/in_ids"])
    assert np.abs(G["subpix/corners"][n0:] - np.rint(G["subpix/corners"][n0:])).max() > 0.05       # the recovered corners did move off the pixel grid


def test_numpy_restatement_edge_cases():
    none = np.zeros((0, 4, 2), np.float32)
    # nothing detected / nothing rejected: inputs come back untouched
    c, i, r, rec = refine_np.refine_detected_markers(G["frame"], DIC, G["board_ids"], G["board_obj"], none, np.zeros(0, np.int32), G["rejected"])
    assert len(i) == 0 and len(rec) == 0 and np.array_equal(r, G["rejected"])
    c, i, r, rec = refine_np.refine_detected_markers(G["frame"], DIC, G["board_ids"], G["board_obj"], G["corners"], G["ids"], none)
    assert np.array_equal(i, G["ids"]) and len(rec) == 0
    # a board that shares no id with the detections: no homography, nothing recovered
    c, i, r, rec = refine_np.refine_detected_markers(G["frame"], DIC, G["board_ids"] + 100, G["board_obj"], G["corners"], G["ids"], G["rejected"])
    assert np.array_equal(i, G["ids"]) and len(rec) == 0
    # homography / pose pieces against closed forms
    H = np.array([[1.1, 0.02, 5.0], [-0.03, 0.95, -2.0], [1e-4, -2e-4, 1.0]])
    src = np.random.default_rng(0).uniform(0, 100, (12, 2))
    assert np.allclose(refine_np.find_homography(src, refine_np.perspective_transform(src, H)), H, atol=1e-9)
    rv, tv = np.array([0.3, -0.2, 0.1]), np.array([0.05, -0.02, 0.6])
    obj = np.c_[np.random.default_rng(1).uniform(-0.1, 0.1, (16, 2)), np.zeros(16)]
    img = refine_np.project_points(obj, rv, tv, G["K"], G["D"])
    r2, t2 = refine_np.solve_pnp_planar(obj, img, G["K"], G["D"])
    assert np.allclose(r2, rv, atol=1e-8) and np.allclose(t2, tv, atol=1e-8)


def test_product_board_geometry_vs_restatement():
    """board_core.h (the product's host-side fits) against the numpy restatement on the golden detections, and against closed forms"""
    from hostemu import emu
    ids = [int(i) for i in G["ids"]]
    bids = [int(i) for i in G["board_ids"]]
    obj = np.concatenate([G["board_obj"][bids.index(i)] for i in ids]).astype(np.float64)
    img = G["corners"].reshape(-1, 2).astype(np.float64)
    ok, H = emu.homography(obj[:, :2], img)
    assert ok and np.abs(refine_np.perspective_transform(obj[:, :2], H) - refine_np.perspective_transform(obj[:, :2], refine_np.find_homography(obj[:, :2], img))).max() < 1e-6
    rc, r, t = emu.board_pose(G["K"], G["D"], obj, img)
    r2, t2 = refine_np.solve_pnp_planar(obj, img, G["K"], G["D"])
    assert rc == 0
    assert np.abs(refine_np.project_points(obj, r, t, G["K"], G["D"]) - refine_np.project_points(obj, r2, t2, G["K"], G["D"])).max() < 1e-6
    # a tilted, offset board plane: exact recovery from exact projections
    rng = np.random.default_rng(4)
    R0 = refine_np.rodrigues(np.array([0.4, 0.3, -0.2]))
    pl = np.c_[rng.uniform(-0.2, 0.2, (20, 2)), np.zeros(20)] @ R0.T + np.array([0.3, -0.1, 0.7])
    rv, tv = np.array([-0.25, 0.5, 0.15]), np.array([-0.2, 0.1, 1.1])
    rc, r, t = emu.board_pose(G["K"], G["D"], pl, refine_np.project_points(pl, rv, tv, G["K"], G["D"]))
    assert rc == 0 and np.allclose(r, rv, atol=1e-7) and np.allclose(t, tv, atol=1e-7)
    # in general position: the DLT start, exact recovery as well; fewer than 6 such points are refused (cv2 throws)
    pl[::3, 2] += 0.05
    rc, r, t = emu.board_pose(G["K"], G["D"], pl, refine_np.project_points(pl, rv, tv, G["K"], G["D"]))
    assert rc == 0 and np.allclose(r, rv, atol=1e-7) and np.allclose(t, tv, atol=1e-7)
    assert emu.board_pose(G["K"], G["D"], pl[:5], refine_np.project_points(pl[:5], rv, tv, G["K"], G["D"]))[0] == 2


NP = np.load(os.path.join(os.path.dirname(__file__), "golden", "refine_nonplanar.npz"))
NP_CASES = [(str(n), tuple(p)) for n, p in zip(NP["cases"], NP["case_params"])]


def test_pose_start_planar_or_dlt_vs_cv2_solvepnp():
    """cv2.solvePnP(ITERATIVE) on points in general position and on nearly planar ones (tests/golden/refine_nonplanar.npz): the numpy
    restatement and the product's board_core.h reach the same reprojection as cv2"""
    from hostemu import emu
    kinds = set()
    for i in range(int(NP["n_pnp"])):
        obj, img = NP["pnp/%d/obj" % i], NP["pnp/%d/img" % i]
        want = refine_np.project_points(obj, NP["pnp/%d/rvec" % i], NP["pnp/%d/tvec" % i], G["K"], G["D"])
        kinds.add(bool(refine_np.is_planar(obj)))
        r, t = refine_np.solve_pnp(obj, img, G["K"], G["D"])
        assert np.abs(refine_np.project_points(obj, r, t, G["K"], G["D"]) - want).max() < 1e-4, i
        rc, r, t = emu.board_pose(G["K"], G["D"], obj, img)
        assert rc == 0 and np.abs(refine_np.project_points(obj, r, t, G["K"], G["D"]) - want).max() < 1e-4, i
    assert kinds == {True, False}


def same_as_cv2_nonplanar(name, c, i, r, rec):
    assert np.array_equal(i, NP[name + "/ids"]) and np.array_equal(c, NP[name + "/corners"])
    assert np.array_equal(r, NP[name + "/rejected"]) and np.array_equal(rec, NP[name + "/recovered"])


@pytest.mark.parametrize("name,prm", NP_CASES)
def test_numpy_restatement_nonplanar_board_vs_cv2(name, prm):
    rep, ecr, orders = prm
    assert not refine_np.is_planar(NP["board_obj"].reshape(-1, 3))
    out = refine_np.refine_detected_markers(G["frame"], DIC, G["board_ids"], NP["board_obj"], G["corners"], G["ids"], G["rejected"], K=G["K"], D=G["D"],
                                            min_rep_distance=rep, error_correction_rate=ecr, check_all_orders=bool(orders))
    same_as_cv2_nonplanar(name, *out)


# ---- the product through the C ABI ----
@pytest.fixture(scope="module")
def aruco():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from aruco_slam_b200 import aruco as A
    return A


@pytest.mark.gpu
@pytest.mark.parametrize("name,prm", CASES)
def test_gpu_refine_detected_markers_vs_cv2(aruco, name, prm):
    rep, ecr, orders, cam = prm
    det = aruco.ArucoDetector(DIC, aruco.DetectorParameters(), max_shape=G["frame"].shape, max_batch=1)
    c, ids, rej = det.detectMarkers(G["frame"])
    board = aruco.Board(G["board_obj"], G["board_ids"])
    out = det.refineDetectedMarkers(G["frame"], board, c, ids, rej, cameraMatrix=G["K"] if cam else None, distCoeffs=G["D"] if cam else None,
                                    refineParams=aruco.RefineParameters(rep, ecr, bool(orders)))
    oc, oi, orj, rec = out
    same_as_cv2(name, np.array(oc, np.float32).reshape(-1, 4, 2), np.asarray(oi, np.int32).ravel(), np.array(orj, np.float32).reshape(-1, 4, 2),
                np.zeros(0, np.int32) if rec is None else np.asarray(rec, np.int32).ravel())
    det.close()


@pytest.mark.gpu
def test_gpu_refine_detected_markers_subpix(aruco):
    det = aruco.ArucoDetector(DIC, aruco.DetectorParameters(cornerRefinementMethod=1), max_shape=G["frame"].shape, max_batch=1)
    c, ids, rej = det.detectMarkers(G["frame"])
    assert np.array_equal(ids.ravel(), G["subpix/in_ids"]) and np.abs(np.array(c).reshape(-1, 4, 2) - G["subpix/in_corners"]).max() < 0.05
    oc, oi, orj, rec = det.refineDetectedMarkers(G["frame"], aruco.Board(G["board_obj"], G["board_ids"]), c, ids, rej)
    assert np.array_equal(oi.ravel(), G["subpix/ids"]) and np.array_equal(rec.ravel(), G["subpix/recovered"])
    assert np.array_equal(np.array(orj).reshape(-1, 4, 2), G["subpix/rejected"])
    assert np.abs(np.array(oc).reshape(-1, 4, 2) - G["subpix/corners"]).max() < 0.05
    det.close()


@pytest.mark.gpu
def test_gpu_refine_detected_markers_edges(aruco):
    det = aruco.ArucoDetector(DIC, aruco.DetectorParameters(), max_shape=G["frame"].shape, max_batch=1)
    c, ids, rej = det.detectMarkers(G["frame"])
    board = aruco.Board(G["board_obj"], G["board_ids"])
    # bgr frame, same answer
    bgr = np.repeat(G["frame"][..., None], 3, axis=2)
    oc, oi, orj, rec = det.refineDetectedMarkers(bgr, board, c, ids, rej)
    assert np.array_equal(np.asarray(oi).ravel(), G["h_default/ids"]) and np.array_equal(np.asarray(rec).ravel(), G["h_default/recovered"])
    # nothing rejected / nothing detected: untouched, recovered None (cv2 returns without writing)
    oc, oi, orj, rec = det.refineDetectedMarkers(G["frame"], board, c, ids, ())
    assert rec is None and np.array_equal(np.asarray(oi).ravel(), G["ids"])
    oc, oi, orj, rec = det.refineDetectedMarkers(G["frame"], board, (), None, rej)
    assert rec is None and len(orj) == len(rej)
    det.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,prm", NP_CASES)
def test_gpu_refine_detected_markers_nonplanar_board_vs_cv2(aruco, name, prm):
    """a board in general position seen through a camera: the pose starts from the DLT, as cv2's does"""
    rep, ecr, orders = prm
    det = aruco.ArucoDetector(DIC, aruco.DetectorParameters(), max_shape=G["frame"].shape, max_batch=1)
    c, ids, rej = det.detectMarkers(G["frame"])
    oc, oi, orj, rec = det.refineDetectedMarkers(G["frame"], aruco.Board(NP["board_obj"], G["board_ids"]), c, ids, rej, cameraMatrix=G["K"], distCoeffs=G["D"],
                                                 refineParams=aruco.RefineParameters(rep, ecr, bool(orders)))
    same_as_cv2_nonplanar(name, np.array(oc, np.float32).reshape(-1, 4, 2), np.asarray(oi, np.int32).ravel(), np.array(orj, np.float32).reshape(-1, 4, 2),
                          np.zeros(0, np.int32) if rec is None else np.asarray(rec, np.int32).ravel())
    det.close()
