"""CPU-only check of the product's kernel *logic* (aruco_slam_b200/csrc/core.h, frame_logic.h,
pose_core.h) compiled for the host as a single lane (tests/hostemu) against the oracle and the
cv2 golden vectors.  The CUDA kernels themselves are exercised by the `-m gpu` tests; this
tier catches algorithmic slips without a GPU."""
import os
import sys

import numpy as np
import pytest

from conftest import golden, golden_names

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostemu"))
import emu  # noqa: E402
from aruco_slam_b200 import dictionaries as D, synth  # noqa: E402


@pytest.mark.parametrize("name", golden_names("detect_") + golden_names("stages_"))
def test_detect_logic_matches_golden(name):
    g = golden(name)
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    r = emu.detect(g["frame"], dic)
    assert r["status"] == 0
    assert np.array_equal(r["ids"], g["ids"])
    assert np.array_equal(r["corners"], g["corners"])
    assert np.array_equal(r["rejected"], g["rejected"])
    if "n_contours" in g.files and min(g["frame"].shape) >= 67:
        assert np.array_equal(r["n_contours"], g["n_contours"])


@pytest.mark.parametrize("name", golden_names("inverted_"))
def test_detect_inverted_logic_matches_golden(name):
    """detectInvertedMarker = true through the product's kernel-logic headers (group walk reversed, flipped bit matrix)"""
    g = golden(name)
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    r = emu.detect(g["frame"], dic, detect_inverted=True)
    assert r["status"] == 0
    assert np.array_equal(r["ids"], g["ids"])
    assert np.array_equal(r["corners"], g["corners"])
    assert np.array_equal(r["rejected"], g["rejected"])


@pytest.mark.parametrize("name", golden_names("stages_"))
def test_contour_points_match_golden(name):
    g = golden(name)
    H, W = g["frame"].shape
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    mn, mx = int(0.03 * max(W, H)), int(4.0 * max(W, H))
    for si in range(3):
        r = emu.detect(g["frame"], dic, dbg_scale=si)
        offs, pts = g["cont_offs%d" % si], g["cont_pts%d" % si]
        lens = np.diff(offs)
        keep = [i for i, n in enumerate(lens) if mn <= n <= mx]
        assert r["n_contours"][si] == len(lens)
        assert np.array_equal(r["kept_len"], lens[keep])
        want = np.concatenate([pts[offs[i]:offs[i + 1]] for i in keep]) if keep else np.zeros((0, 2), np.int16)
        assert np.array_equal(r["kept_pts"], want)


@pytest.mark.parametrize("R", [1, 4, 8, 32])
@pytest.mark.parametrize("p", [0.2, 0.5, 0.8])
def test_contours_random_masks_vs_oracle(oracle, p, R):
    """border graph (anchors every R rows / columns -> segments -> cycles -> emit) vs the oracle's
    findContours restatement on random masks: contour count, kept lengths, every point"""
    rng = np.random.default_rng(int(p * 10))
    m = ((rng.random((97, 131)) < p) * 255).astype(np.uint8)
    dic = D.getPredefinedDictionary(0)
    r = emu.detect(m, dic, masks=np.stack([m, m, m]), dbg_scale=0, anchor_R=R)
    assert r["status"] >= 0             # negative = the emulation's own consistency checks; 3 = candidate capacity (irrelevant here)
    cs = oracle.find_contours(m)
    mn, mx = int(0.03 * 131), 4 * 131
    kept = [c for c in cs if mn <= len(c) <= mx]
    assert r["n_contours"][0] == len(cs)
    assert np.array_equal(r["kept_len"], [len(c) for c in kept])
    assert np.array_equal(r["kept_pts"], np.concatenate(kept).astype(np.int16))


def test_pose_logic_matches_golden():
    g = golden("pose")
    K, Dd = g["K"], g["D"]
    wr = wt = 0.0
    for row in g["rows"]:
        L, use_d = float(row[0]), int(row[1])
        r, t = emu.pose(row[2:10].astype(np.float32), K, Dd if use_d else np.zeros(5), L)
        wr = max(wr, np.abs(r[0] - row[10:13]).max())
        wt = max(wt, np.abs(t[0] - row[13:16]).max())
    assert wr < 1e-4 and wt < 1e-4          # north_star tolerance
    assert wr < 1e-5 and wt < 1e-5, (wr, wt)


def test_observation_logic_matches_oracle(oracle):
    g = golden("pose")
    K, Dd = g["K"], g["D"]
    sp = oracle.slam_params(useful_distance_threshold=3.0)
    kept = 0
    for i, row in enumerate(g["rows"][::7]):
        corners = row[2:10].astype(np.float32)
        dist = Dd if int(row[1]) else np.zeros(5)
        sp.marker_length = float(row[0])
        obs = oracle.make_observations(corners.reshape(1, 4, 2), [i], row[10:13].reshape(1, 3), row[13:16].reshape(1, 3), K, dist, sp)
        ok, o = emu.observation(corners, i, row[10:13], row[13:16], K, dist, sp)
        assert ok == (len(obs) == 1)
        if ok:
            kept += 1
            assert abs(o[0] - obs[0].x) < 1e-12 and abs(o[1] - obs[0].y) < 1e-12 and abs(o[2] - obs[0].theta) < 1e-12
            assert np.abs(o[3:] - np.array(obs[0].cov[:])).max() < 1e-12
    assert kept > 5


def test_pose_logic_detected_markers():
    g = golden("pose_detected")
    rows = g["rows"]
    r, t = emu.pose(rows[:, 2:10].astype(np.float32), g["K"], g["D"], 0.27)
    assert np.abs(t - rows[:, 13:16]).max() < 1e-4                                                   # m
    assert max(synth.rvec_distance(a, b) for a, b in zip(r, rows[:, 10:13])) < 1e-4                  # rad (rotation)
    assert np.abs(r - rows[:, 10:13]).max() < 5e-4                                                   # same rvec branch


def test_otsu_restatement_matches_textbook_loop():
    """The division-free, range-restricted Otsu (prefix sums + reciprocal + fma correction) must pick the same
    threshold as OpenCV's sequential loop, including the plateau / tie cases of two-level patches."""
    rng = np.random.default_rng(11)
    cases = []
    for _ in range(400):                                    # random sparse / dense histograms of 1024 samples
        k = int(rng.integers(1, 200))
        bins = rng.choice(256, size=k, replace=False)
        cases.append(np.bincount(rng.choice(bins, size=1024), minlength=256))
    for a, b in ((0, 255), (30, 220), (100, 101), (5, 6), (254, 255), (0, 1)):      # two spikes (plateau between them)
        for na in (1, 17, 512, 1023):
            h = np.zeros(256, int); h[a] = na; h[b] = 1024 - na; cases.append(h)
    h = np.zeros(256, int); h[77] = 576; cases.append(h)                            # flat patch
    cases.append(np.full(256, 4))                                                   # uniform
    for _ in range(100):                                    # blurred two-level patches like a warped marker
        v = np.clip(np.concatenate([rng.normal(40, 6, 500), rng.normal(210, 9, 400), rng.uniform(40, 210, 124)]), 0, 255).astype(int)
        cases.append(np.bincount(v, minlength=256))
    for n in (576, 784, 5184):                              # patches whose pixel count is not a power of two (1 / n inexact): sparse histograms with long empty runs
        for _ in range(150):
            k = int(rng.integers(2, 40))
            bins = rng.choice(256, size=k, replace=False)
            cases.append(np.bincount(rng.choice(bins, size=n), minlength=256))
    for h in cases:
        new, seq = emu.otsu(h)
        assert new == seq, (new, seq, np.nonzero(h)[0][:8])


def test_approx_closed_vs_oracle_on_many_contours(oracle):
    """the product's approxPolyDP(closed) (32-bit proxies, two trackers, first-maximum rule) against the oracle's
    restatement of OpenCV 4.13 on every contour of noisy / blobby masks, at the detector's epsilon and a finer one"""
    rng = np.random.default_rng(21)
    masks = []
    for p in (0.35, 0.55, 0.7):
        masks.append(((rng.random((120, 160)) < p) * 255).astype(np.uint8))
    yy, xx = np.mgrid[0:200, 0:260]
    blob = np.zeros((200, 260), np.uint8)
    for _ in range(25):                                 # overlapping rotated rectangles and discs: long borders with straight runs
        cx, cy, a, b, th = rng.uniform(20, 240), rng.uniform(20, 180), rng.uniform(5, 45), rng.uniform(5, 45), rng.uniform(0, 3.14)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th); v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        blob |= ((np.abs(u) < a) & (np.abs(v) < b)).astype(np.uint8) * 255 if rng.random() < 0.7 else ((u * u + v * v) < a * a).astype(np.uint8) * 255
    masks.append(blob)
    checked = quads = 0
    for m in masks:
        for c in oracle.find_contours(m):
            n = len(c)
            if n < 8:
                continue
            for rate in (0.03, 0.01):
                want = oracle.approx_poly_dp(c, n * rate)
                got = emu.approx(c, n * rate)
                if got is None:                         # the product stops once more than 8 vertices are certain
                    assert len(want) > 4
                else:
                    if len(want) <= 4 or len(got) <= 4:
                        assert np.array_equal(got, want), (n, rate)
                    quads += len(want) == 4
                checked += 1
    assert checked > 300 and quads > 10
