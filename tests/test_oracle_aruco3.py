"""Pin the oracle's ArUco3 path (useAruco3Detection: image pyramid, reduced segmentation image, identification in the pyramid level
that suits each candidate, corner refinement up the pyramid) to what cv2 4.13.0 returns: tests/golden/aruco3.npz, written by
tools/make_golden_aruco3.py.  No GPU, no cv2."""
import numpy as np
import pytest

from conftest import golden
from aruco_slam_b200 import dictionaries as D

A3 = golden("aruco3")
SUBPIX_TOL = 0.05        # accepted corners go through cornerSubPix (float32 accumulations): same tolerance as the SUBPIX tests


def a3_params(oracle, g, fixture, case):
    i = list(A3["cases"]).index(case)
    extra = {"r015_inv": {"detectInvertedMarker": 1}, "r020_contour": {"cornerRefinementMethod": 2}}.get(case, {})
    return oracle.default_params(useAruco3Detection=1, minSideLengthCanonicalImg=int(A3["sides"][i]),
                                 minMarkerLengthRatioOriginalImg=float(A3["%s/%s/ratio" % (fixture, case)]), **extra)


def test_pyr_down_equals_cv2(oracle):
    for i in range(int(A3["n_pyr"])):
        assert np.array_equal(oracle.pyr_down(A3["pyr/%d/src" % i]), A3["pyr/%d/dst" % i]), i


def test_resize_linear_equals_cv2(oracle):
    for i in range(int(A3["n_resize"])):
        dst = A3["resize/%d/dst" % i]
        assert np.array_equal(oracle.resize_linear(A3["resize/%d/src" % i], dst.shape[1], dst.shape[0]), dst), i


@pytest.mark.parametrize("fixture", [str(f) for f in A3["fixtures"]])
def test_detect_aruco3(oracle, fixture):
    g = golden(fixture)
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    n_markers = 0
    for case in [str(c) for c in A3["cases"]]:
        key = "%s/%s" % (fixture, case)
        c, ids, rej = oracle.detect(g["frame"], dic, a3_params(oracle, g, fixture, case))
        assert np.array_equal(ids, A3[key + "/ids"]), key
        assert np.array_equal(rej, A3[key + "/rejected"]), key              # segmentation-image coordinates, bit exact
        if len(ids):
            assert np.abs(c - A3[key + "/corners"]).max() <= SUBPIX_TOL, key
        n_markers += len(ids)
    assert n_markers > 0 or fixture == "detect_blank"


def test_aruco3_off_leaves_the_two_parameters_without_effect(oracle):
    g = golden("detect_vga_4x4_s2")
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    c, ids, rej = oracle.detect(g["frame"], dic, oracle.default_params(minSideLengthCanonicalImg=64, minMarkerLengthRatioOriginalImg=0.05))
    assert np.array_equal(ids, g["ids"]) and np.array_equal(c, g["corners"]) and np.array_equal(rej, g["rejected"])


# ---- the product's ArUco3 logic (pyr_core.h, frame_logic.h) compiled for the host: tests/hostemu ----
@pytest.fixture(scope="module")
def emu():
    from hostemu import emu as E
    E.lib()
    return E


def test_product_pyr_down_and_resize_equal_cv2(emu):
    for i in range(int(A3["n_pyr"])):
        assert np.array_equal(emu.pyr_down(A3["pyr/%d/src" % i]), A3["pyr/%d/dst" % i]), i
    for i in range(int(A3["n_resize"])):
        dst = A3["resize/%d/dst" % i]
        assert np.array_equal(emu.resize_linear(A3["resize/%d/src" % i], dst.shape[1], dst.shape[0]), dst), i


def test_product_pyramid_equals_oracle_on_frames(emu, oracle):
    g = golden("detect_vga_4x4_s2")["frame"]
    cur_e = cur_o = g
    for _ in range(4):
        cur_e, cur_o = emu.pyr_down(cur_e), oracle.pyr_down(cur_o)
        assert np.array_equal(cur_e, cur_o)
    for dw, dh in ((457, 343), (320, 240), (213, 160), (639, 479)):
        assert np.array_equal(emu.resize_linear(g, dw, dh), oracle.resize_linear(g, dw, dh))


@pytest.mark.parametrize("fixture", [str(f) for f in A3["fixtures"] if "1080p" not in str(f) and "720p" not in str(f)])
def test_product_logic_aruco3_vs_cv2(emu, fixture):
    """level choice, scaled identification, perimeter gate of the product headers: ids and rejected equal cv2's; the accepted
    corners before the refinement chain are the refined ones to within the refinement's reach"""
    g = golden(fixture)
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    H, W = g["frame"].shape
    for ci, case in enumerate([str(c) for c in A3["cases"]]):
        key = "%s/%s" % (fixture, case)
        out = emu.detect_aruco3(g["frame"], dic, int(A3["sides"][ci]), float(A3[key + "/ratio"]), detect_inverted=(case == "r015_inv"))
        assert out is not None and out["status"] == 0, key
        assert np.array_equal(out["ids"], A3[key + "/ids"]), key
        assert np.array_equal(out["rejected"], A3[key + "/rejected"]), key
        if len(out["ids"]):
            up = out["corners"] * (W / out["plan"][0])
            assert np.abs(up - A3[key + "/corners"]).max() <= 2.0 * (W / out["plan"][0]) + 3.0, key


def test_product_logic_aruco3_odd_sizes_vs_oracle(emu, oracle):
    from aruco_slam_b200 import synth
    fr = synth.render_config("C1", 5).image
    dic = D.getPredefinedDictionary(0)
    plans = set()
    for W, H in ((639, 479), (637, 475), (333, 250), (201, 199), (64, 48), (40, 30)):
        crop = np.ascontiguousarray(fr[:H, :W])
        for ratio, side in ((0.0, 32), (0.02, 32), (0.05, 16), (0.01, 8)):
            oc, oi, orj = oracle.detect(crop, dic, oracle.default_params(useAruco3Detection=1, minSideLengthCanonicalImg=side, minMarkerLengthRatioOriginalImg=ratio))
            out = emu.detect_aruco3(crop, dic, side, ratio)
            assert out is not None and np.array_equal(out["ids"], oi) and np.array_equal(out["rejected"], orj), (W, H, ratio, side)
            plans.add(out["plan"][2:])
    assert (5, 2) in plans and (0, 0) in plans          # a refinement chain over two levels, and a pyramid of one image


def test_interior_ranges_equal_the_per_group_test(emu):
    """k_pyr_down walks the interior groups by ranges (pyr_interior_gx_end / pyr_interior_y_end): the same set as pyr_down_is_interior4"""
    import ctypes as C
    f = emu.lib().emu_pyr_ranges_ok
    for W in list(range(1, 70)) + [639, 640, 641, 1919, 1920]:
        for H in (1, 2, 3, 4, 5, 9, 34, 479):
            assert f(W, H) == 1, (W, H)


def test_plans_the_product_refuses(emu):
    """aruco3_plan says no where cv2 would index past its pyramid (the level closest to the segmentation image does not exist) or
    where the pyramid needs more than 12 images; the create / detect calls turn that into B2A_ERR_INVALID / B2A_ERR_UNSUPPORTED"""
    dic = D.getPredefinedDictionary(0)
    small = np.full((48, 64), 120, np.uint8)
    assert emu.detect_aruco3(small, dic, 8, 2.0) is None               # segmentation image 4 x 3: its level 4 lies past a pyramid of 3 images
    assert emu.detect_aruco3(small, dic, 8, 0.05) is not None
    wide = np.full((4, 20000), 120, np.uint8)
    assert emu.detect_aruco3(wide, dic, 1, 0.0) is not None            # log2(80000) / 2 = 8 levels: fine
