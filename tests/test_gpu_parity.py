"""-m gpu: the CUDA path (through the C ABI, include/b2aruco.h) against the CPU oracle and the
cv2 golden vectors.  Bars (BASELINE.json north_star): ids and threshold masks bit-exact;
corners <= 0.05 px; rvec <= 1e-4 rad, tvec <= 1e-4 m; accepted / rejected lists in the
reference's order.  Without sub-pixel refinement corners are integer valued, so they are
compared for equality."""
import zlib

import numpy as np
import pytest

from conftest import golden, golden_names
from aruco_slam_b200 import dictionaries as D, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def aruco():
    from aruco_slam_b200 import aruco as A
    return A


def _detector(A, dic, shape, batch=1, **prm):
    return A.ArucoDetector(dic, A.DetectorParameters(**prm), max_shape=shape, max_batch=batch)


@pytest.mark.parametrize("name", golden_names("stages_"))
def test_stage_taps_vs_golden(aruco, oracle, name):
    g = golden(name)
    gray = g["frame"]
    H, W = gray.shape
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    det = _detector(aruco, dic, gray.shape)
    # A1 + A2 through the colour path
    bgr = synth.gray_to_bgr(gray, 7)
    gg, masks = det.debug_threshold(bgr)
    assert zlib.crc32(gg[0].tobytes()) == int(g["bgr_gray_crc"])
    # A2 on the gray frame: masks bit-exact
    gg, masks = det.debug_threshold(gray)
    assert np.array_equal(gg[0], gray)
    for si in range(3):
        assert np.array_equal(np.packbits(masks[0, si] > 0), g["mask%d" % si]), si
    # A3a: contour count, kept contours (order, lengths, points)
    counts, kept, lens, pts = det.debug_contours(gray, pts_cap=W * H)
    mn, mx = int(0.03 * max(W, H)), int(4.0 * max(W, H))
    for si in range(3):
        offs, gp = g["cont_offs%d" % si], g["cont_pts%d" % si]
        L = np.diff(offs)
        keep = [i for i, n in enumerate(L) if mn <= n <= mx]
        assert counts[0, si] == len(L)
        assert kept[0, si] == len(keep)
        assert np.array_equal(lens[0, si, :len(keep)], L[keep])
        want = np.concatenate([gp[offs[i]:offs[i + 1]] for i in keep]) if keep else np.zeros((0, 2), np.int16)
        assert np.array_equal(pts[0, si, :len(want)], want)
    # A3/A4 candidates vs oracle
    _, _, _, dbg = oracle.detect(gray, dic, debug=True)
    cand = det.debug_candidates(gray)[0]
    assert np.array_equal(cand, dbg["cand"])
    # final
    c, ids, rej = det.detectMarkers(gray)
    assert np.array_equal(np.array(c, np.float32).reshape(-1, 4, 2), g["corners"])
    assert np.array_equal(ids.ravel() if ids is not None else np.zeros(0, np.int32), g["ids"])
    assert np.array_equal(np.array(rej, np.float32).reshape(-1, 4, 2), g["rejected"])
    c, ids, rej = det.detectMarkers(bgr)
    assert np.array_equal(np.array(c, np.float32).reshape(-1, 4, 2), g["bgr_corners"])
    assert np.array_equal(ids.ravel(), g["bgr_ids"])
    assert np.array_equal(np.array(rej, np.float32).reshape(-1, 4, 2), g["bgr_rejected"])
    det.close()


@pytest.mark.parametrize("name", golden_names("detect_"))
def test_detect_vs_golden(aruco, name):
    g = golden(name)
    gray = g["frame"]
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    det = _detector(aruco, dic, gray.shape)
    r = det.detect_batch(gray)
    assert np.array_equal(r.ids[0], g["ids"])                 # bit exact, same order
    assert np.array_equal(r.corners[0], g["corners"])
    assert np.array_equal(r.rejected[0], g["rejected"])
    _, masks = det.debug_threshold(gray)
    for si in range(3):
        assert zlib.crc32(masks[0, si].tobytes()) == int(g["mask_crc"][si])
    det.close()
    det = _detector(aruco, dic, gray.shape, cornerRefinementMethod=1)
    r = det.detect_batch(gray)
    assert np.array_equal(r.ids[0], g["subpix_ids"])
    if len(r.ids[0]):
        assert np.abs(r.corners[0] - g["subpix_corners"]).max() < 0.05     # north_star: 0.05 px
        assert np.abs(r.corners[0] - g["subpix_corners"]).max() < 1e-2
    det.close()


@pytest.mark.parametrize("name", golden_names("inverted_"))
def test_detect_inverted_vs_golden(aruco, name):
    """detectInvertedMarker = true on the GPU path against cv2's output (bit exact, same order)"""
    g = golden(name)
    gray = g["frame"]
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    det = _detector(aruco, dic, gray.shape, detectInvertedMarker=1)
    r = det.detect_batch(gray)
    assert np.array_equal(r.ids[0], g["ids"])
    assert np.array_equal(r.corners[0], g["corners"])
    assert np.array_equal(r.rejected[0], g["rejected"])
    det.close()


def test_legacy_free_functions(aruco):
    g = golden("detect_vga_4x4_s2")
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    corners, ids, rejected = aruco.detectMarkers(g["frame"], dic)
    assert isinstance(corners, tuple) and corners[0].shape == (1, 4, 2) and corners[0].dtype == np.float32
    assert ids.shape == (len(corners), 1) and ids.dtype == np.int32
    assert np.array_equal(ids.ravel(), g["ids"])
    blank = golden("detect_blank")
    corners, ids, rejected = aruco.detectMarkers(blank["frame"], dic)
    assert corners == () and ids is None                      # cv2: zero detections -> ids None
    with pytest.raises(aruco.B2AError):
        aruco.detectMarkers(np.zeros((0, 0), np.uint8), dic)    # cv::Exception in the reference
    with pytest.raises(aruco.B2AError):
        aruco.estimatePoseSingleMarkers(np.zeros((1, 4, 2), np.float32), 0.0, np.eye(3), np.zeros(5))


def test_library_dictionaries_match_package(aruco):
    for did in (D.DICT_4X4_50, D.DICT_5X5_100, D.DICT_6X6_250, D.DICT_7X7_1000, D.DICT_ARUCO_ORIGINAL, D.DICT_APRILTAG_36H11):
        a, b = aruco.library_dictionary(did), D.getPredefinedDictionary(did)
        assert a.marker_size == b.marker_size and a.max_correction_bits == b.max_correction_bits
        assert np.array_equal(a.table, b.table)


def test_known_answer_generated_markers(aruco):
    """generateImageMarker(id) -> detect -> id (SURVEY 8c known-answer identity), all four rotations."""
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    det = _detector(aruco, dic, (400, 400))
    for mid in (0, 17, 123, 249):
        for rot in range(4):
            img = np.full((400, 400), 255, np.uint8)
            m = np.rot90(dic.marker_image(mid, 30), rot)
            img[80:80 + m.shape[0], 80:80 + m.shape[1]] = m
            r = det.detect_batch(img)
            assert r.ids[0].tolist() == [mid], (mid, rot, r.ids[0])
    det.close()


def test_batch_vs_oracle_1080p(aruco, oracle):
    """config C2 frames in one batch: every frame identical to the oracle (ids, corners, rejected, order)."""
    B = 4
    frames = synth.render_batch("C2", B, base_seed=100)
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    det = _detector(aruco, dic, frames.shape[1:], batch=B)
    K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]])
    Dist = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
    r = det.detect_pose_batch(frames, 0.27, K, Dist)
    for b in range(B):
        oc, oi, orj = oracle.detect(frames[b], dic)
        assert np.array_equal(r.ids[b], oi) and np.array_equal(r.corners[b], oc) and np.array_equal(r.rejected[b], orj)
        orv, otv = oracle.estimate_pose_single_markers(oc, 0.27, K, Dist)
        assert np.abs(r.tvecs[b] - otv).max() < 1e-4
        # same rotation within 1e-4 rad (the rotation *vector* may sit on the other side of a half turn)
        assert max(synth.rvec_distance(a, c) for a, c in zip(r.rvecs[b], orv)) < 1e-4
    # idempotence: a second call on the same handle returns the same result
    r2 = det.detect_pose_batch(frames, 0.27, K, Dist)
    for b in range(B):
        assert np.array_equal(r.ids[b], r2.ids[b]) and np.array_equal(r.corners[b], r2.corners[b])
        assert np.array_equal(r.rvecs[b], r2.rvecs[b])
    # a sub-batch and a smaller frame on the same handle
    r3 = det.detect_batch(frames[1:2])
    assert np.array_equal(r3.ids[0], r.ids[1])
    small = synth.render_config("C1", 3).image
    det2 = _detector(aruco, D.getPredefinedDictionary(0), frames.shape[1:], batch=B)
    oc, oi, orj = oracle.detect(small, D.getPredefinedDictionary(0))
    r4 = det2.detect_batch(small)
    assert np.array_equal(r4.ids[0], oi) and np.array_equal(r4.corners[0], oc) and np.array_equal(r4.rejected[0], orj)
    det.close(); det2.close()


def test_noisy_4k_frame_vs_oracle(aruco, oracle):
    """config C3 (4K, noise + blur), one frame."""
    fr = synth.render_config("C3", 0).image
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    det = _detector(aruco, dic, fr.shape)
    r = det.detect_batch(fr)
    oc, oi, orj = oracle.detect(fr, dic)
    assert np.array_equal(r.ids[0], oi) and np.array_equal(r.corners[0], oc) and np.array_equal(r.rejected[0], orj)
    det.close()


def test_pose_vs_golden(aruco):
    g = golden("pose")
    K, Dd = g["K"], g["D"]
    det = _detector(aruco, D.getPredefinedDictionary(0), (64, 64))
    rows = g["rows"]
    for use_d in (0, 1):
        for L in (0.27, 0.1):
            sel = rows[(rows[:, 1] == use_d) & (np.abs(rows[:, 0] - L) < 1e-9)]
            rv, tv = det.estimatePoseSingleMarkers(sel[:, 2:10].astype(np.float32).reshape(-1, 4, 2), L, K, Dd if use_d else np.zeros(5))
            assert rv.shape == (len(sel), 1, 3)
            assert np.abs(rv.reshape(-1, 3) - sel[:, 10:13]).max() < 1e-4      # rad
            assert np.abs(tv.reshape(-1, 3) - sel[:, 13:16]).max() < 1e-4      # m
    gd = golden("pose_detected")
    rows = gd["rows"]
    rv, tv = det.estimatePoseSingleMarkers(rows[:, 2:10].astype(np.float32).reshape(-1, 4, 2), 0.27, gd["K"], gd["D"])
    assert np.abs(tv.reshape(-1, 3) - rows[:, 13:16]).max() < 1e-4
    assert max(synth.rvec_distance(a, b) for a, b in zip(rv.reshape(-1, 3), rows[:, 10:13])) < 1e-4
    assert np.abs(rv.reshape(-1, 3) - rows[:, 10:13]).max() < 5e-4      # same rotation-vector branch as cv2
    det.close()


def test_threshold_edge_shapes(aruco, oracle):
    """ragged sizes: widths not multiples of 16/32/128, tiny frames, non-default window list."""
    rng = np.random.default_rng(3)
    dic = D.getPredefinedDictionary(0)
    for (H, W) in ((31, 33), (64, 129), (95, 257), (130, 64), (7, 300), (50, 324), (26, 644), (100, 12), (49, 963), (241, 5)):
        img = rng.integers(0, 256, (H, W)).astype(np.uint8)
        det = _detector(aruco, dic, (H, W))
        _, masks = det.debug_threshold(img)
        for si, k in enumerate((3, 13, 23)):
            assert np.array_equal(masks[0, si], oracle.adaptive_threshold(img, k, 7.0)), (H, W, k)
        det.close()
    img = rng.integers(0, 256, (120, 200)).astype(np.uint8)
    det = _detector(aruco, dic, img.shape, adaptiveThreshWinSizeMin=5, adaptiveThreshWinSizeMax=29, adaptiveThreshWinSizeStep=8, adaptiveThreshConstant=3.0)
    assert det.num_scales == 4
    _, masks = det.debug_threshold(img)
    for si, k in enumerate((5, 13, 21, 29)):
        assert np.array_equal(masks[0, si], oracle.adaptive_threshold(img, k, 3.0)), k
    det.close()


def test_bgr_ingest_fused_into_threshold(aruco, oracle):
    """bgr8 frames: the marching threshold kernel converts on the fly and writes the gray plane (widths that are multiples of 4);
    other widths take the separate conversion pass.  Gray plane and masks against the oracle on random colour frames, batches,
    multi-chunk work items and ragged last strips; detections on a rendered frame equal those of its gray version."""
    rng = np.random.default_rng(21)
    dic = D.getPredefinedDictionary(0)
    for (H, W, B) in ((31, 36, 1), (130, 64, 2), (50, 324, 1), (26, 644, 3), (241, 8, 1), (700, 1000, 2), (95, 257, 1), (64, 130, 2)):
        img = rng.integers(0, 256, (B, H, W, 3)).astype(np.uint8)
        img[0, : H // 2] = (img[0, : H // 2] // 64) * 64 + 63            # flat areas too
        det = _detector(aruco, dic, (H, W), batch=B)
        gg, masks = det.debug_threshold(img if B > 1 else img[0])
        for b in range(B):
            want = oracle.bgr2gray(img[b])
            assert np.array_equal(gg[b], want), (H, W, b)
            for si, k in enumerate((3, 13, 23)):
                assert np.array_equal(masks[b, si], oracle.adaptive_threshold(want, k, 7.0)), (H, W, b, k)
        det.close()
    fr = synth.render_config("C2", 5).image
    bgr = synth.gray_to_bgr(fr, 3)
    det = _detector(aruco, D.getPredefinedDictionary(10), fr.shape, batch=2)
    r = det.detect_batch(np.stack([bgr, bgr[:, ::-1].copy()]))
    oc, oi, orj = oracle.detect(bgr, D.getPredefinedDictionary(10))
    assert len(oi) >= 25 and np.array_equal(r.ids[0], oi) and np.array_equal(r.corners[0], oc) and np.array_equal(r.rejected[0], orj)
    oc, oi, orj = oracle.detect(bgr[:, ::-1].copy(), D.getPredefinedDictionary(10))
    assert np.array_equal(r.ids[1], oi) and np.array_equal(r.corners[1], oc) and np.array_equal(r.rejected[1], orj)
    det.close()


def test_graph_replay_equals_plain_launches(aruco, oracle):
    """calls of up to 4 frames replay a CUDA graph captured per (batch, shape, camera): different frames, batch sizes, shapes and
    cameras through one handle give what plain launches give (and the oracle), and per-stage times exist only for plain launches"""
    dic = D.getPredefinedDictionary(0)
    frs = [synth.render_config("C1", s).image for s in range(10, 16)]
    K1 = np.array([[600.0, 0, 320], [0, 600.0, 240], [0, 0, 1]])
    K2 = np.array([[450.0, 0, 300], [0, 470.0, 250], [0, 0, 1]])
    det = _detector(aruco, dic, frs[0].shape, batch=4)
    plain = _detector(aruco, dic, frs[0].shape, batch=4)
    plain.set_graph(False)
    seq = [(frs[0], K1), (frs[1], K1), (np.stack(frs[2:5]), K1), (frs[5], K2), (frs[0][:400, :600].copy(), K2), (frs[1], K1), (np.stack(frs[1:5]), K2),
           (synth.gray_to_bgr(frs[2], 1), K1), (frs[3], K1)]
    for img, K in seq:
        a = det.detect_pose_batch(img, 0.05, K, np.array([0.05, -0.02, 0.001, 0.0, 0.0]))
        b = plain.detect_pose_batch(img, 0.05, K, np.array([0.05, -0.02, 0.001, 0.0, 0.0]))
        for f in range(len(a.ids)):
            assert np.array_equal(a.ids[f], b.ids[f]) and np.array_equal(a.corners[f], b.corners[f]) and np.array_equal(a.rejected[f], b.rejected[f])
            assert np.array_equal(a.rvecs[f], b.rvecs[f]) and np.array_equal(a.tvecs[f], b.tvecs[f])
        one = img if img.ndim == 2 or img.shape[-1] == 3 and img.ndim == 3 else img[0]
        oc, oi, _ = oracle.detect(one, dic)
        assert np.array_equal(a.ids[0], oi) and np.array_equal(a.corners[0], oc)
    assert det.last_launch_count() == plain.last_launch_count() > 10
    assert sum(det.last_stage_times().values()) == 0 and sum(plain.last_stage_times().values()) > 0
    det.close(); plain.close()


def test_threshold_device_frames_pitches(aruco, oracle):
    """frames already in device memory: an unaligned row pitch takes the tiled kernel, a padded (4-byte aligned) pitch and a
    frame stride take the marching kernel; both must give the oracle's masks and the same detections as host frames"""
    import ctypes as C
    import torch
    from aruco_slam_b200 import _lib
    rng = np.random.default_rng(11)
    dic = D.getPredefinedDictionary(0)
    H, W = 95, 257
    imgs = rng.integers(0, 256, (2, H, W)).astype(np.uint8)
    det = _detector(aruco, dic, (H, W), batch=2)
    want = [[oracle.adaptive_threshold(im, k, 7.0) for k in (3, 13, 23)] for im in imgs]
    for pitch, fstride in ((W, W * H), (260, 260 * H + 64)):
        buf = torch.zeros(2 * fstride + 64, dtype=torch.uint8, device="cuda")
        for b in range(2):
            view = buf[b * fstride: b * fstride + pitch * H].view(H, pitch)
            view[:, :W] = torch.from_numpy(imgs[b]).cuda()
            view[:, W:] = 0xEE                                   # padding bytes must never be read as pixels
        fr = aruco.ArucoDetector.frames_device(buf.data_ptr(), 2, H, W, 1, pitch, fstride)
        gray = np.zeros((2, H, W), np.uint8)
        masks = np.zeros((2, 3, H, W), np.uint8)
        _lib.check(_lib.lib().b2a_debug_threshold(det._h, C.byref(fr), gray.ctypes.data, masks.ctypes.data))
        assert np.array_equal(gray, imgs)
        for b in range(2):
            for si in range(3):
                assert np.array_equal(masks[b, si], want[b][si]), (pitch, b, si)
    det.close()


def test_contours_random_masks(aruco, oracle):
    """Bernoulli masks through the contour stages (threshold of a two-level image reproduces the mask)."""
    rng = np.random.default_rng(9)
    dic = D.getPredefinedDictionary(0)
    for p in (0.3, 0.5, 0.7):
        H, W = 97, 131
        # a frame whose k=3 mask is the Bernoulli pattern is hard to construct; instead compare
        # against the oracle's contours of the mask the GPU itself produced (bit-exact masks are
        # checked above)
        img = rng.integers(0, 256, (H, W)).astype(np.uint8)
        det = _detector(aruco, dic, (H, W))
        _, masks = det.debug_threshold(img)
        counts, kept, lens, pts = det.debug_contours(img, pts_cap=2 * W * H)
        mn, mx = int(0.03 * max(W, H)), int(4.0 * max(W, H))
        for si in range(3):
            cs = oracle.find_contours(masks[0, si])
            keep = [c for c in cs if mn <= len(c) <= mx]
            assert counts[0, si] == len(cs)
            assert kept[0, si] == len(keep)
            assert np.array_equal(lens[0, si, :len(keep)], [len(c) for c in keep])
            want = np.concatenate(keep).astype(np.int16) if keep else np.zeros((0, 2), np.int16)
            assert np.array_equal(pts[0, si, :len(want)], want)
        det.close()


def test_results_do_not_depend_on_streams_or_anchor_grid(aruco, oracle):
    """sub-batch streams and uneven host splits are scheduling only: identical detections for 1, 4 and 8 streams,
    device-resident and host frames"""
    B = 12
    frames = synth.render_batch("C2", B, base_seed=40)
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    det = _detector(aruco, dic, frames.shape[1:], batch=B)
    ref = None
    for n in (1, 4, 8):
        det.set_streams(n)
        r = det.detect_batch(frames)
        if ref is None:
            ref = r
            for b in (0, B - 1):
                oc, oi, orj = oracle.detect(frames[b], dic)
                assert np.array_equal(r.ids[b], oi) and np.array_equal(r.corners[b], oc) and np.array_equal(r.rejected[b], orj)
        else:
            for b in range(B):
                assert np.array_equal(r.ids[b], ref.ids[b]) and np.array_equal(r.corners[b], ref.corners[b])
                assert np.array_equal(r.rejected[b], ref.rejected[b])
    det.close()


def test_border_starting_at_the_first_pixel(aruco, oracle):
    """a border whose first point is pixel (0,0) has start key 0 (regression: it used to collide with the sort padding)"""
    g = np.full((96, 128), 255, np.uint8)
    g[0:2, 0:60] = 0                        # a thin dark L hugging the top-left corner: every window size thresholds it
    g[0:60, 0:2] = 0                        # to a region whose outer border starts at pixel (0,0)
    g[30:32, 40:110] = 0; g[70:72, 40:110] = 0; g[30:72, 40:42] = 0; g[30:72, 108:110] = 0      # and a thin ring elsewhere
    assert oracle.adaptive_threshold(g, 13, 7.0)[0, 0] == 255
    dic = D.getPredefinedDictionary(0)
    det = _detector(aruco, dic, g.shape)
    for _ in range(2):                      # twice: the second call sees stale scratch from the first
        counts, n_kept, kept_len, pts = det.debug_contours(g)
        for si, k in enumerate((3, 13, 23)):
            mk = oracle.adaptive_threshold(g, k, 7.0)
            cs = oracle.find_contours(mk)
            mn, mx = int(0.03 * 128), 4 * 128
            kept = [c for c in cs if mn <= len(c) <= mx]
            assert counts[0, si] == len(cs)
            assert n_kept[0, si] == len(kept)
            assert np.array_equal(kept_len[0, si, :len(kept)], [len(c) for c in kept])
            want = np.concatenate(kept).astype(np.int16) if kept else np.zeros((0, 2), np.int16)
            assert np.array_equal(pts[0, si, :len(want)], want)
    det.close()


def test_capacity_overflow_is_reported(aruco):
    """B2A_ERR_CAPACITY (status 3): more markers than max_markers, more quad candidates than max_candidates.  The call
    fails loudly, the per-frame status names the frame, and the handle stays usable."""
    from aruco_slam_b200._lib import B2AError
    frames = np.stack([synth.render_config("C2", 7).image, synth.background(1920, 1080).astype(np.uint8)])
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    det = aruco.ArucoDetector(dic, max_shape=frames.shape[1:], max_batch=2, max_markers=8)
    with pytest.raises(B2AError) as e:
        det.detect_batch(frames)
    assert e.value.code == 3
    r = det.detect_batch(frames[1:])                   # the blank frame alone is fine on the same handle
    assert len(r.ids[0]) == 0
    det.close()
    det = aruco.ArucoDetector(dic, max_shape=frames.shape[1:], max_batch=2, max_candidates=32)
    with pytest.raises(B2AError) as e:
        det.detect_batch(frames)
    assert e.value.code == 3
    det.close()


def test_speckled_1080p_batch_fits_at_default_streams(aruco, oracle):
    """heavy sensor noise at 1080p (about 0.5 M border anchors per frame) at the default sub-batch count and at eight
    sub-batches: every frame keeps its full share of the border-graph arrays whatever the cut (no B2A_ERR_CAPACITY),
    and the results equal the oracle's"""
    B = 3
    rng = np.random.default_rng(11)
    frames = np.stack([np.clip(synth.render_config("C2", 40 + b).image.astype(np.float64) + rng.normal(0, 22.0, (1080, 1920)), 0, 255).astype(np.uint8)
                       for b in range(B)])
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    det = _detector(aruco, dic, frames.shape[1:], batch=8)        # a handle sized for more frames than the call brings
    for streams in (0, 8, 1):
        det.set_streams(streams)
        r = det.detect_batch(frames)
        for b in range(B):
            oc, oi, orj = oracle.detect(frames[b], dic)
            assert np.array_equal(r.ids[b], oi) and np.array_equal(r.corners[b], oc) and np.array_equal(r.rejected[b], orj), (streams, b)
    det.close()


def test_bench_batch_all_frames(aruco, oracle):
    """the exact batch bench.py times by default (C2, seeds 0-31), every frame against the oracle, host and device frames"""
    import torch
    B = 32
    frames = synth.render_batch("C2", B, base_seed=0)
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    det = _detector(aruco, dic, frames.shape[1:], batch=B)
    K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]])
    Dist = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
    r = det.detect_pose_batch(frames, 0.27, K, Dist)
    dfr = torch.from_numpy(frames).cuda()
    cam = aruco._camera(K, Dist, 0.27)
    r2 = det._collect(det.detect_raw(aruco.ArucoDetector.frames_device(dfr.data_ptr(), B, 1080, 1920), cam), True)
    total = 0
    for b in range(B):
        oc, oi, orj = oracle.detect(frames[b], dic)
        assert np.array_equal(r.ids[b], oi) and np.array_equal(r.corners[b], oc) and np.array_equal(r.rejected[b], orj), b
        assert np.array_equal(r2.ids[b], oi) and np.array_equal(r2.corners[b], oc) and np.array_equal(r2.rejected[b], orj), b
        orv, otv = oracle.estimate_pose_single_markers(oc, 0.27, K, Dist)
        assert np.abs(r.tvecs[b] - otv).max() < 1e-4 and max(synth.rvec_distance(a, c) for a, c in zip(r.rvecs[b], orv)) < 1e-4
        assert np.array_equal(r.rvecs[b], r2.rvecs[b]) and np.array_equal(r.tvecs[b], r2.tvecs[b])
        total += len(oi)
    assert total > 28 * B
    det.close()


def test_submit_wait_matches_synchronous_calls(aruco, oracle):
    """b2a_detect_pose_submit / _wait: two batches in flight on one handle give the results of the synchronous call, in
    order; the misuse cases fail loudly"""
    from aruco_slam_b200._lib import B2AError
    B = 4
    batches = [synth.render_batch("C2", B, base_seed=200 + 10 * k) for k in range(5)]
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    det = _detector(aruco, dic, batches[0].shape[1:], batch=B)
    K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]])
    Dist = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
    want = [det.detect_pose_batch(b, 0.27, K, Dist) for b in batches]
    got = list(det.detect_pose_stream(batches, 0.27, K, Dist))
    assert len(got) == len(want)
    for w, g in zip(want, got):
        for b in range(B):
            assert np.array_equal(w.ids[b], g.ids[b]) and np.array_equal(w.corners[b], g.corners[b]) and np.array_equal(w.rejected[b], g.rejected[b])
            assert np.array_equal(w.rvecs[b], g.rvecs[b]) and np.array_equal(w.tvecs[b], g.tvecs[b])
    oc, oi, _ = oracle.detect(batches[3][1], dic)
    assert np.array_equal(got[3].ids[1], oi) and np.array_equal(got[3].corners[1], oc)
    # misuse: a third submit, a synchronous call while a batch is in flight, a stale ticket
    cam = aruco._camera(K, Dist, 0.27)
    fr, keep = det._frames_host(batches[0])
    t0 = det.submit_raw(fr, cam)
    t1 = det.submit_raw(fr, cam)
    with pytest.raises(B2AError):
        det.submit_raw(fr, cam)
    with pytest.raises(B2AError):
        det.detect_pose_batch(batches[0], 0.27, K, Dist)
    det.wait_raw(t0)
    with pytest.raises(B2AError):
        det.wait_raw(t0)
    r = det._collect(det.wait_raw(t1), True)
    assert np.array_equal(r.ids[0], want[0].ids[0])
    det.detect_pose_batch(batches[0], 0.27, K, Dist)          # and the synchronous call works again
    det.close()


def test_multi_device_detector_gathers_in_frame_order(aruco, oracle):
    """b2a_multi_*: one handle and one host thread per listed device, contiguous blocks, detections gathered on the host in frame
    order.  On a one-GPU box the two handles share device 0; with more devices visible they spread."""
    import torch
    n_dev = torch.cuda.device_count()
    devices = tuple(range(min(n_dev, 4))) if n_dev >= 2 else (0, 0)
    B = 7                                               # not a multiple of the device count: ragged blocks
    frames = synth.render_batch("C2", B, base_seed=300)
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]])
    Dist = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
    md = aruco.MultiDetector(dic, devices=devices, max_shape=frames.shape[1:], max_batch=B)
    r = md.detect_pose_batch(frames, 0.27, K, Dist)
    single = _detector(aruco, dic, frames.shape[1:], batch=B)
    r1 = single.detect_pose_batch(frames, 0.27, K, Dist)
    for b in range(B):
        oc, oi, orj = oracle.detect(frames[b], dic)
        assert np.array_equal(r.ids[b], oi) and np.array_equal(r.corners[b], oc) and np.array_equal(r.rejected[b], orj), b
        assert np.array_equal(r.rvecs[b], r1.rvecs[b]) and np.array_equal(r.tvecs[b], r1.tvecs[b])
    r2 = md.detect_pose_batch(frames[:1], 0.27, K, Dist)          # fewer frames than devices: empty blocks
    assert np.array_equal(r2.ids[0], r.ids[0])
    r3 = md.detect_pose_batch(frames[2:5])                        # detect only
    assert r3.rvecs is None and np.array_equal(r3.ids[1], r.ids[3])
    md.close(); single.close()


@pytest.mark.parametrize("name", golden_names("draw_"))
def test_draw_detected_markers_vs_cv2(aruco, name):
    """b2a_draw_detected_markers against cv2.aruco.drawDetectedMarkers overlays, every pixel (with ids, without, custom colour)"""
    from test_draw import draw_case
    g = golden(name)
    img, want = draw_case(g)
    det = _detector(aruco, D.getPredefinedDictionary(0), img.shape[:2])
    assert np.array_equal(det.drawDetectedMarkers(img.copy(), g["corners"], g["ids"]), want["ids"])
    assert np.array_equal(det.drawDetectedMarkers(img.copy(), g["corners"]), want["no_ids"])
    assert np.array_equal(det.drawDetectedMarkers(img.copy(), g["corners"], g["ids"], tuple(int(v) for v in g["colour"])), want["colour"])
    # what the reference does per frame: detect, then draw on a copy (aruco_slam.cpp:313-319)
    c, ids, _ = det.detectMarkers(img) if int(g["ids"].max()) < 50 else (None, None, None)
    if ids is not None and sorted(ids.ravel().tolist()) == sorted(g["ids"].tolist()):
        assert np.array_equal(det.drawDetectedMarkers(img.copy(), np.array(c), ids), want["ids"])
    det.close()


def test_pack_detections_round_trip(aruco):
    """b2a_pack_detections: the compact record of a call (what a multi-process gather moves) unpacks to the same detections"""
    frames = synth.render_batch("C2", 3, base_seed=400)
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    det = _detector(aruco, dic, frames.shape[1:], batch=3)
    K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]])
    cam = aruco._camera(K, np.zeros(5), 0.27)
    fr, keep = det._frames_host(frames)
    raw = det.detect_raw(fr, cam)
    want = det._collect(raw, True)
    buf = np.zeros(1 << 16, np.uint8)
    n = aruco.pack_detections(raw, buf)
    got = aruco.unpack_detections(buf[:n])
    assert 0 < n < 20000
    for b in range(3):
        assert np.array_equal(got.ids[b], want.ids[b]) and np.array_equal(got.corners[b], want.corners[b]) and np.array_equal(got.rejected[b], want.rejected[b])
        assert np.array_equal(got.rvecs[b], want.rvecs[b]) and np.array_equal(got.tvecs[b], want.tvecs[b])
    from aruco_slam_b200._lib import B2AError
    with pytest.raises(B2AError):
        aruco.pack_detections(raw, np.zeros(64, np.uint8))
    det.close()


# ---- ArUco3 (useAruco3Detection) ------------------------------------------------------------------------------------------
_A3 = golden("aruco3")
_A3_EXTRA = {"r015_inv": {"detectInvertedMarker": 1}, "r020_contour": {"cornerRefinementMethod": 2}}


@pytest.mark.parametrize("fixture", [str(f) for f in _A3["fixtures"]])
def test_aruco3_vs_golden(aruco, fixture):
    """useAruco3Detection on the GPU path against cv2 4.13 (tests/golden/aruco3.npz): ids and the rejected quads (which cv2 leaves in
    the coordinates of the reduced image) bit exact and in order, accepted corners (refined up the pyramid) <= 0.05 px"""
    g = golden(fixture)
    gray = g["frame"]
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    for ci, case in enumerate([str(c) for c in _A3["cases"]]):
        key = "%s/%s" % (fixture, case)
        det = _detector(aruco, dic, gray.shape, useAruco3Detection=1, minSideLengthCanonicalImg=int(_A3["sides"][ci]),
                        minMarkerLengthRatioOriginalImg=float(_A3[key + "/ratio"]), **_A3_EXTRA.get(case, {}))
        r = det.detect_batch(gray)
        assert np.array_equal(r.ids[0], _A3[key + "/ids"]), key
        assert np.array_equal(r.rejected[0], _A3[key + "/rejected"]), key
        if len(r.ids[0]):
            assert np.abs(r.corners[0] - _A3[key + "/corners"]).max() < 0.05, key
        det.close()


def test_aruco3_batch_colour_pose_vs_oracle(aruco, oracle):
    """a batch of colour 1080p frames through the sub-batch streams and, one frame at a time, through the graph replay: every frame
    equal to the oracle's ArUco3 path; the poses follow from the refined corners"""
    B = 6
    gray = synth.render_batch("C2", B, base_seed=300)
    bgr = np.stack([synth.gray_to_bgr(f, 3 + i) for i, f in enumerate(gray)])
    dic = D.getPredefinedDictionary(D.DICT_6X6_250)
    K = np.array([[1400.0, 0, 960], [0, 1400.0, 540], [0, 0, 1]])
    Dist = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
    for ratio, side in ((0.02, 32), (0.0, 32), (32 / 1920, 32)):
        prm = dict(useAruco3Detection=1, minSideLengthCanonicalImg=side, minMarkerLengthRatioOriginalImg=ratio)
        det = _detector(aruco, dic, gray.shape[1:], batch=B, **prm)
        r = det.detect_pose_batch(bgr, 0.27, K, Dist)
        n = 0
        for b in range(B):
            oc, oi, orj = oracle.detect(bgr[b], dic, oracle.default_params(**prm))
            assert np.array_equal(r.ids[b], oi) and np.array_equal(r.rejected[b], orj), (ratio, b)
            if len(oi):
                assert np.abs(r.corners[b] - oc).max() < 0.05, (ratio, b)
                orv, otv = oracle.estimate_pose_single_markers(r.corners[b], 0.27, K, Dist)
                assert np.abs(r.tvecs[b] - otv).max() < 1e-4
                assert max(synth.rvec_distance(a, c) for a, c in zip(r.rvecs[b], orv)) < 1e-4
            n += len(oi)
        assert n > 3 * B
        # single frames (graph replay, gray input) give the same as the batch
        for b in (0, B - 1):
            r1 = det.detect_batch(gray[b])
            oc, oi, orj = oracle.detect(gray[b], dic, oracle.default_params(**prm))
            assert np.array_equal(r1.ids[0], oi) and np.array_equal(r1.rejected[0], orj)
            if len(oi):
                assert np.abs(r1.corners[0] - oc).max() < 0.05
            r2 = det.detect_batch(gray[b])
            assert np.array_equal(r1.corners[0], r2.corners[0]) and np.array_equal(r1.ids[0], r2.ids[0])
        det.close()


def test_aruco3_parameter_checks(aruco):
    from aruco_slam_b200._lib import B2AError
    dic = D.getPredefinedDictionary(0)
    with pytest.raises(B2AError):
        _detector(aruco, dic, (480, 640), useAruco3Detection=1, minSideLengthCanonicalImg=0)
    with pytest.raises(B2AError):
        _detector(aruco, dic, (480, 640), useAruco3Detection=1, minMarkerLengthRatioOriginalImg=-0.1)
    # the two parameters do nothing while the mode is off
    g = golden("detect_vga_4x4_s2")
    det = _detector(aruco, dic, g["frame"].shape, minSideLengthCanonicalImg=64, minMarkerLengthRatioOriginalImg=0.05)
    r = det.detect_batch(g["frame"])
    assert np.array_equal(r.ids[0], g["ids"]) and np.array_equal(r.corners[0], g["corners"]) and np.array_equal(r.rejected[0], g["rejected"])
    det.close()


def test_aruco3_odd_sizes_and_two_level_chain(aruco, oracle):
    """frame sizes that are not multiples of 4 (byte path of the pyramid's first level, ragged last groups of every level), tiny frames
    (a pyramid of one image) and parameters whose refinement climbs two levels: equal to the oracle's ArUco3 path"""
    fr = synth.render_config("C1", 5).image
    dic = D.getPredefinedDictionary(0)
    found = 0
    for W, H in ((639, 479), (637, 475), (333, 250), (64, 48)):
        crop = np.ascontiguousarray(fr[:H, :W])
        for ratio, side in ((0.0, 32), (0.05, 16), (0.01, 8)):
            prm = dict(useAruco3Detection=1, minSideLengthCanonicalImg=side, minMarkerLengthRatioOriginalImg=ratio)
            det = _detector(aruco, dic, crop.shape, **prm)
            r = det.detect_batch(crop)
            oc, oi, orj = oracle.detect(crop, dic, oracle.default_params(**prm))
            assert np.array_equal(r.ids[0], oi) and np.array_equal(r.rejected[0], orj), (W, H, ratio, side)
            if len(oi):
                assert np.abs(r.corners[0] - oc).max() < 0.05, (W, H, ratio, side)
            found += len(oi)
            det.close()
    assert found >= 10
