"""The arithmetic of the marching threshold kernel (k_threshold_march, detect_kernels.cuh) restated in NumPy
and checked against the CPU oracle, so that the CPU test tier covers the identities the kernel relies on:

  * column prefixes kept as two packed u16 lanes per 32-bit word (even / odd bytes): a vertical window sum is one
    plain 32-bit subtraction of two prefixes -- no borrow between the lanes while (Hs + 22) * 255 < 2^16;
  * replicate border = clamp the word index + a byte permute;
  * horizontal box sums by sliding (add the entering, subtract the leaving u16 element), started from the
    threshold bias, and the test  S - k^2 g - c_k >= 0  with  c_k = k^2 C - (k^2 - 1) / 2   (exact integers)
    <=>  g - ((2 S + k^2) div (2 k^2)) <= -C   (cv2.adaptiveThreshold MEAN_C / BINARY_INV, reference
    src/aruco_slam.cpp:313 via detectMarkers).
The GPU tier (test_gpu_parity.py::test_threshold_edge_shapes, test_stage_taps_vs_golden) checks the kernel itself.
"""
import numpy as np
import pytest

WT, HALO, RC = 320, 12, 24          # TM_WT, TM_HALO, TM_RC


def march_item(gray, X0, Y0, Hs, radii, Cfloor):
    """masks of one work item (strip X0.., rows Y0..Y0+Hs) computed the way the kernel computes them"""
    H, W = gray.shape
    rows = min(Hs, H - Y0)
    ncol = (WT + 2 * HALO) // 4                                   # V-phase threads
    wmax = (W - 1) >> 2
    padW = 4 * (wmax + 1)
    gp = np.zeros((H, padW), np.uint8)
    gp[:, :W] = gray
    gp[:, W:] = 0xEE                                              # bytes past the row end are never selected
    words = gp.view("<u4")                                        # [H][wmax + 1]
    wcol = (X0 - HALO) // 4 + np.arange(ncol)
    wc = np.clip(wcol, 0, wmax)
    bj = np.clip(4 * wcol[:, None] + np.arange(4)[None, :], 0, W - 1) - 4 * wc[:, None]          # byte picked for column j
    ytop = Y0 - (HALO - 1)
    n_raw = ((rows + RC - 1) // RC) * RC + 22
    ce = np.zeros(ncol, np.uint32)
    co = np.zeros(ncol, np.uint32)
    Ce = np.zeros((n_raw + 1, ncol), np.uint32)                   # Ce[j + 1] = prefix through raw index j
    Co = np.zeros((n_raw + 1, ncol), np.uint32)
    for i in range(n_raw):
        gy = min(max(ytop + i, 0), H - 1)
        w = words[gy, wc]
        byte = lambda sel: (w >> (8 * sel).astype(np.uint32)) & np.uint32(0xFF)
        ce = ce + (byte(bj[:, 0]) | (byte(bj[:, 2]) << np.uint32(16)))
        co = co + (byte(bj[:, 1]) | (byte(bj[:, 3]) << np.uint32(16)))
        Ce[i + 1], Co[i + 1] = ce, co
    assert int(max(ce.max() & 0xFFFF, ce.max() >> 16)) < 65536
    out = np.zeros((len(radii), rows, WT), bool)
    for k, r in enumerate(radii):
        K2 = (2 * r + 1) ** 2
        bias = -(K2 * Cfloor - (K2 - 1) // 2)
        for m in range(rows):
            ve = Ce[m + r + 12] - Ce[m + 11 - r]                  # packed: no borrow between the lanes
            vo = Co[m + r + 12] - Co[m + 11 - r]
            V = np.empty(4 * ncol, np.int64)                      # plane row: column q of the strip's V
            V[0::4] = ve & 0xFFFF; V[2::4] = ve >> 16; V[1::4] = vo & 0xFFFF; V[3::4] = vo >> 16
            gy = Y0 + m
            for seg in range(WT // 64):
                x0 = 64 * seg
                if X0 + x0 >= W:
                    continue
                S = bias + int(V[x0 + HALO - r: x0 + HALO + r + 1].sum())
                for i in range(64):
                    x = X0 + x0 + i
                    if x < W:
                        out[k, m, x0 + i] = (S - K2 * int(gray[gy, x])) >= 0
                    S += int(V[x0 + i + HALO + 1 + r]) - int(V[x0 + i + HALO - r])
    return out


@pytest.mark.parametrize("shape,Hs,Cf", [((60, 324), 48, 7), ((30, 12), 24, 7), ((75, 650), 72, 3), ((49, 37), 24, -2)])
def test_march_arithmetic_matches_oracle(oracle, shape, Hs, Cf):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    H, W = shape
    gray = rng.integers(0, 256, (H, W)).astype(np.uint8)
    gray[: H // 3, : W // 2] = 255                                # saturated block: the largest sums and prefixes
    gray[H // 2:, W // 2:] = np.where(rng.random((H - H // 2, W - W // 2)) < 0.5, 0, 255)
    radii = (1, 6, 11)
    want = [oracle.adaptive_threshold(gray, 2 * r + 1, float(Cf)) > 0 for r in radii]
    for Y0 in range(0, H, Hs):
        for X0 in range(0, W, WT):
            got = march_item(gray, X0, Y0, Hs, radii, Cf)
            rows = got.shape[1]
            cols = min(WT, W - X0)
            for k in range(3):
                assert np.array_equal(got[k, :, :cols], want[k][Y0:Y0 + rows, X0:X0 + cols]), (shape, X0, Y0, radii[k])


def test_packed_prefix_bound():
    """the tallest work item the host may choose keeps every packed prefix below 2^16"""
    TM_MAX_HS = 216
    assert (TM_MAX_HS + 22) * 255 < 1 << 16
    assert TM_MAX_HS % RC == 0
