"""cv::aruco::drawDetectedMarkers (reference src/aruco_slam.cpp:319; getMarkedImg, aruco_slam.h:152): the product's drawing logic
(aruco_slam_b200/csrc/draw_core.h compiled for the host) against overlays written by cv2 4.13.0 (tests/golden/draw_*.npz,
tools/make_overlay_tables.py) -- every pixel, with ids, without ids, and with a custom border colour.  The CUDA path is checked
against the same files in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import golden, golden_names
from aruco_slam_b200 import synth


def draw_case(g):
    gray = g["gray"]
    img = gray if int(g["bgr_seed"]) < 0 else synth.gray_to_bgr(gray, int(g["bgr_seed"]))
    want = {}
    for key in ("ids", "no_ids", "colour"):
        w = img.copy().ravel()
        w[g["idx_" + key]] = g["val_" + key]
        want[key] = w.reshape(img.shape)
    return img, want


@pytest.mark.parametrize("name", golden_names("draw_"))
def test_draw_logic_vs_cv2(name):
    from hostemu import emu
    g = golden(name)
    img, want = draw_case(g)
    assert np.array_equal(emu.draw(img, g["corners"], g["ids"]), want["ids"])
    assert np.array_equal(emu.draw(img, g["corners"]), want["no_ids"])
    assert np.array_equal(emu.draw(img, g["corners"], g["ids"], tuple(int(v) for v in g["colour"])), want["colour"])
    assert len(g["idx_ids"]) > 500
