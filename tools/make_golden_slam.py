#!/usr/bin/env python
"""Write tests/golden/slam_*.npz and map_txt.npz by RUNNING THE REFERENCE ITSELF (authoring container only).

oracle/_ref/libarucoslam_ref.so is /root/reference/src/aruco_slam.cpp + src/map_loader.cpp compiled unmodified
(oracle/Makefile, stand-in Eigen / OpenCV / ROS headers in oracle/ref_stubs/).  Its five OpenCV calls are routed into
the cv2 4.13.0 wheel here (oracle.ref.use_cv2_hooks), so what is recorded is "reference source + real OpenCV":
observations in the order the reference's priority queue pops them, mu / Sigma after every frame, the stationary-gate
bookkeeping, toRosPose, the MarkerArrays, MapLoader's records.  The CPU oracle (oracle/orc_*.c), the host emulation of the
product headers and the CUDA path are pinned to these files (tests/test_slam_golden.py, tests/test_gpu_slam.py).

Run:  python tools/make_golden_slam.py      (needs /root/reference and cv2)
"""
import os
import sys
import tempfile

import numpy as np
import cv2

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from aruco_slam_b200 import synth, dictionaries as D  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
PROV = "reference src/aruco_slam.cpp + src/map_loader.cpp compiled unmodified (oracle/_ref) + cv2 %s hooks; tools/make_golden_slam.py" % cv2.__version__

# /root/reference/default.yaml:10-20
K_REF = np.array([[525.2866213437447, 0, 472.85738972861157], [0, 525.2178123117577, 264.77181506420266], [0, 0, 1]])
D_REF = np.array([0.04160142651680036, -0.04771035303381654, -0.0032638387781624705, -0.003985120051161831, 0.01110263483766991])
L = 0.27                                                        # parameters.yaml:17


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, provenance=np.array(PROV), **kw)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def record_frame(kw, f, r, img):
    """taps around one addImage: the queue (private getObservations), then the full call"""
    oid, oidx, oxyt, ocov = r.get_observations(img)
    kw["obs_id_%d" % f], kw["obs_index_%d" % f], kw["obs_xyt_%d" % f], kw["obs_cov_%d" % f] = oid, oidx, oxyt, ocov
    r.add_image(img)
    mu, sg, ids = r.get_state()
    kw["mu_%d" % f], kw["sigma_%d" % f], kw["ids_%d" % f] = mu, sg, ids
    lid, lobs = r.last_observed()
    kw["last_id_%d" % f], kw["last_obs_%d" % f] = lid, lobs
    p, q, c = r.robot_pose()
    kw["pose_%d" % f] = np.concatenate([p, q, c])
    dm = r.markers(1)
    kw["detm_id_%d" % f], kw["detm_pos_%d" % f], kw["detm_q_%d" % f] = dm["id"], dm["position"], dm["orientation"]


def synth_detections(n, rng, K, Dc, far=False, noisy=False):
    """detections whose observation (x, y, theta) is under control: rotation about the camera's Y axis by -theta."""
    ids = rng.choice(60, n, replace=False).astype(np.int32)
    rv, tv, cs = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 4, 2), np.float32)
    h = np.float32(L) / np.float32(2)
    obj = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float32)
    for k in range(n):
        th = rng.uniform(-1, 1)
        rv[k] = [rng.normal(0, 0.05), -th, rng.normal(0, 0.05)]
        tv[k] = [rng.uniform(-1, 1), rng.uniform(-0.3, 0.3), rng.uniform(0.8, 2.7)]
        if far and k % 3 == 1:
            tv[k, 2] = rng.uniform(2.95, 3.3)                    # around the range gate (USEFUL_DISTANCE_THRESHOLD 3)
        p, _ = cv2.projectPoints(obj, rv[k], tv[k], K, Dc)
        sig = 0.15 if not (noisy and k % 2 == 0) else rng.uniform(0.5, 2.5)     # large reprojection error: covariance gate
        cs[k] = (p.reshape(4, 2) + rng.normal(0, sig, (4, 2))).astype(np.float32)
    return cs, ids, rv, tv


def scenario_synth():
    """detections replayed into the reference (no images): gates, duplicates, repeated frames, many new landmarks at once"""
    ref.use_cv2_hooks()
    K = np.array([[600.0, 0, 320], [0, 600.0, 240], [0, 0, 1]])
    Dc = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
    r2c = (0.1, 0.02, 0.3)
    r = ref.RefSlam(r2c_t=r2c)
    r.set_camera(K, Dc)
    rng = np.random.default_rng(2024)
    kw = dict(K=K, D=Dc, r2c_t=np.array(r2c), marker_length=L, useful_distance_threshold=np.float32(3.0))
    img = np.zeros((8, 8), np.uint8)
    # before the first encoder message addImage is ignored (:84-85)
    ref.set_replay(*synth_detections(3, rng, K, Dc))
    r.add_image(img)
    assert r.dim == 3 and not r.is_init
    t = 10.0
    r.add_encoder(1.0, 1.0, t)          # first message only latches the clock (:24-29)
    mu, sg, _ = r.get_state()
    assert r.is_init and not mu.any() and not sg.any()
    enc, last = [], None
    F = 16
    for f in range(F):
        n_enc = 1 + f % 3
        e = []
        for _ in range(n_enc):
            dt = float(rng.uniform(0.02, 0.12))
            t += dt
            wl, wr = (0.0, 0.0) if f in (5, 6) else (float(rng.uniform(-1, 6)), float(rng.uniform(-1, 6)))
            r.add_encoder(wl, wr, t)
            e.append((wl, wr, dt))
        enc.append(e)
        if f in (3, 4, 5, 11, 12) and last is not None:
            det = last                                              # identical detections again: stationary gate (:192-198)
        else:
            det = synth_detections(int(rng.integers(2, 9)), rng, K, Dc, far=(f % 2 == 1), noisy=(f % 4 == 2))
            if f == 8:                                              # the same marker twice in one frame
                cs, ids, rv, tv = det
                det = (np.concatenate([cs, cs[:1] + 0.25]), np.append(ids, ids[0]), np.concatenate([rv, rv[:1]]), np.concatenate([tv, tv[:1] + 0.01]))
        last = det
        ref.set_replay(*det)
        kw["det_corners_%d" % f], kw["det_ids_%d" % f], kw["det_rvecs_%d" % f], kw["det_tvecs_%d" % f] = det
        kw["enc_%d" % f] = np.array(e)
        record_frame(kw, f, r, img)
    kw["n_frames"] = F
    mm = r.markers(0)
    kw["map_id"], kw["map_pos"], kw["map_q"], kw["map_scale"] = mm["id"], mm["position"], mm["orientation"], mm["scale"]
    ref.set_replay()
    save("slam_synth", **kw)
    counts = [len(kw["obs_id_%d" % f]) for f in range(F)]
    dets = [len(kw["det_ids_%d" % f]) for f in range(F)]
    print("   detections per frame", dets, "observations kept", counts, "final dim", r.dim)
    r.close()


SCENE_MAP = synth.REFERENCE_MAP + (
    (7, 0.27, 3.0, 0.6025, 0.3, 1.5708, -0.0, 0.0), (8, 0.27, 1.0, 0.6025, 0.3, 1.5708, -0.0, 0.0),
    (9, 0.27, 5.10375, -0.75, 0.3, 0.0, -1.5708, 0.0), (10, 0.27, 5.10375, -2.25, 0.3, 0.0, -1.5708, 0.0),
    (11, 0.27, 3.0, -4.09375, 0.3, -1.5708, -0.0, 0.0), (12, 0.27, 1.0, -4.09375, 0.3, -1.5708, -0.0, 0.0),
    (13, 0.27, 5.10375, 0.3, 0.75, 0.0, -1.5708, 0.0), (14, 0.27, 5.10375, -1.1, 0.75, 0.0, -1.5708, 0.0),
    (15, 0.27, 3.5, -2.2, 0.3, -1.5708, -0.0, 0.0), (16, 0.27, 4.3, -2.2, 0.3, -1.5708, -0.0, 0.0), (17, 0.27, 2.7, -2.2, 0.3, -1.5708, -0.0, 0.0),
    (18, 0.27, 3.9, -2.2, 0.8, -1.5708, -0.0, 0.0),
)


def scenario_scene():
    """rendered frames of a map.txt-style room seen from a driving robot, /root/reference/default.yaml intrinsics,
    DICT_ARUCO_ORIGINAL (parameters.yaml:16): detectMarkers / solvePnP / Rodrigues / projectPoints are cv2's."""
    ref.use_cv2_hooks()
    W, H = 960, 540
    r2c = (0.12, 0.0, 0.25)
    r = ref.RefSlam(r2c_t=r2c, markers_dictionary=16)
    r.set_camera(K_REF, D_REF)
    kw = dict(K=K_REF, D=D_REF, r2c_t=np.array(r2c), marker_length=L, useful_distance_threshold=np.float32(3.0), dict_id=16,
              scene_map=np.array(SCENE_MAP, float))
    pose = np.array([2.0, -0.3, 0.0])
    t = 0.0
    r.add_encoder(0.0, 0.0, t)
    # (wl, wr) per frame, 5 encoder messages of 0.1 s between frames
    plan = [(8, 8), (8, 8), (8, 8), (8, 8), (8, 8), (0, 0), (0, 0), (6, 2), (6, 2), (6, 2), (8, 8), (8, 8), (5, 7), (8, 8)]
    frames, truth = [], []
    for f, (wl, wr) in enumerate(plan):
        e = []
        for _ in range(5):
            t += 0.1
            r.add_encoder(float(wl), float(wr), t)
            pose = synth.drive(pose, wl, wr, 0.1)
            e.append((wl, wr, 0.1))
        kw["enc_%d" % f] = np.array(e, float)
        poses = synth.scene_poses(SCENE_MAP, pose, r2c, K_REF, D_REF, W, H)
        fr = synth.render_scene(W, H, 16, K_REF, D_REF, L, poses, seed=f, noise_sigma=1.5 if f % 2 else 0.0, blur_sigma=0.7)
        frames.append(fr.image)
        truth.append(pose.copy())
        # what cv2 itself returns on this frame (the tests feed the same frame to the oracle / the GPU)
        det = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(16), cv2.aruco.DetectorParameters())
        c, ids, _ = det.detectMarkers(fr.image)
        n = 0 if ids is None else len(ids)
        kw["det_ids_%d" % f] = np.zeros(0, np.int32) if ids is None else ids.ravel().astype(np.int32)
        kw["det_corners_%d" % f] = np.array(c, np.float32).reshape(-1, 4, 2)
        h = np.float32(L) / np.float32(2)
        obj = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float32)
        rv, tv = np.zeros((n, 3)), np.zeros((n, 3))
        for k in range(n):
            _, a, b = cv2.solvePnP(obj, kw["det_corners_%d" % f][k].reshape(-1, 1, 2), K_REF, D_REF)
            rv[k], tv[k] = a.ravel(), b.ravel()
        kw["det_rvecs_%d" % f], kw["det_tvecs_%d" % f] = rv, tv
        record_frame(kw, f, r, fr.image)
    kw["frames"] = np.stack(frames)
    kw["truth"] = np.array(truth)
    kw["n_frames"] = len(plan)
    mm = r.markers(0)
    kw["map_id"], kw["map_pos"], kw["map_q"], kw["map_scale"] = mm["id"], mm["position"], mm["orientation"], mm["scale"]
    save("slam_scene", **kw)
    print("   detections per frame", [len(kw["det_ids_%d" % f]) for f in range(len(plan))], "observations kept",
          [len(kw["obs_id_%d" % f]) for f in range(len(plan))], "final dim", r.dim)
    mu, _, ids = r.get_state()
    print("   robot estimate", np.round(mu[:3], 3), "truth", np.round(pose, 3), "landmark ids", ids.tolist())
    r.close()


def c5_detections(mu0, ids, n_obs, step, K, Dc):
    """detections whose observations are the true relative poses of n_obs known landmarks + noise"""
    r = np.random.default_rng(100 + step)
    n_lm = (len(mu0) - 3) // 3
    h = np.float32(L) / np.float32(2)
    obj = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float32)
    cs, rv, tv, oid = [], [], [], []
    for k in r.choice(n_lm, n_obs, replace=False):
        Lk = 3 + 3 * k
        c, s = np.cos(mu0[2]), np.sin(mu0[2])
        dx, dy = mu0[Lk] - mu0[0], mu0[Lk + 1] - mu0[1]
        z = np.array([dx * c + dy * s, -dx * s + dy * c, mu0[Lk + 2] - mu0[2]]) + r.normal(0, 0.02, 3)
        # scale into the range gate: the EKF only needs z; keep |t| < 3 by observing through a unit-less camera frame
        z[:2] *= 0.3
        rvec = np.array([0.0, -z[2], 0.0])
        tvec = np.array([-z[1], 0.05, z[0] if z[0] > 0.4 else 0.4 + abs(z[0])])
        p, _ = cv2.projectPoints(obj, rvec, tvec, K, Dc)
        cs.append((p.reshape(4, 2) + r.normal(0, 0.1, (4, 2))).astype(np.float32))
        rv.append(rvec); tv.append(tvec); oid.append(ids[k])
    return np.array(cs, np.float32), np.array(oid, np.int32), np.array(rv), np.array(tv)


def scenario_c5(n_lm, frames, n_obs, name, full):
    """EKF-only workload: known landmarks, set_state, n_obs corrections per frame (BASELINE config 5 at a size whose
    covariance fits a fixture; the 500-landmark run stores summaries only)"""
    ref.use_cv2_hooks()
    K = np.array([[600.0, 0, 320], [0, 600.0, 240], [0, 0, 1]])
    Dc = np.zeros(5)
    mu0, sigma0, ids = synth.c5_state(n_lm)
    r = ref.RefSlam()
    r.set_camera(K, Dc)
    r.set_state(mu0, sigma0, ids, is_init=True)
    kw = dict(K=K, D=Dc, n_lm=n_lm, n_obs=n_obs, n_frames=frames, marker_length=L, r2c_t=np.zeros(3), useful_distance_threshold=np.float32(3.0))
    img = np.zeros((8, 8), np.uint8)
    for f in range(frames):
        det = c5_detections(mu0, ids, n_obs, f, K, Dc)
        ref.set_replay(*det)
        kw["det_corners_%d" % f], kw["det_ids_%d" % f], kw["det_rvecs_%d" % f], kw["det_tvecs_%d" % f] = det
        oid, oidx, oxyt, ocov = r.get_observations(img)
        kw["obs_id_%d" % f], kw["obs_index_%d" % f], kw["obs_xyt_%d" % f], kw["obs_cov_%d" % f] = oid, oidx, oxyt, ocov
        r.add_image(img)
        mu, sg, _ = r.get_state()
        kw["mu_%d" % f] = mu
        if full:
            kw["sigma_%d" % f] = sg
        else:                                   # summaries of an 18 MB matrix: diagonal, three rows, a strided sample, norms
            kw["sigma_diag_%d" % f] = np.diag(sg).copy()
            kw["sigma_rows_%d" % f] = sg[[0, 1, 2, 3 + 3 * (n_lm // 2), len(mu) - 1]].copy()
            kw["sigma_sample_%d" % f] = sg[::37, ::41].copy()
            kw["sigma_fro_%d" % f] = np.linalg.norm(sg)
            kw["sigma_sum_%d" % f] = sg.sum()
        print("   %s frame %d: %d observations" % (name, f, len(oid)))
    ref.set_replay()
    save(name, **kw)
    r.close()


def scenario_map():
    """MapLoader on the reference's own map/map.txt and on the loader's edge cases (map_loader.cpp:20-84)"""
    kw = {}
    m = ref.map_load(os.path.join(ref.REFERENCE, "map", "map.txt"))
    for k, v in m.items():
        kw["ref_" + k] = v
    kw["ref_text"] = np.array(open(os.path.join(ref.REFERENCE, "map", "map.txt")).read())
    cases = {
        "comments_blank": "# a comment\n\n   \n3 0.2 1 2\n  # indented comment\n4 0.3 -1.5 2.5 0.4 0.1 0.2 0.3\n",
        "malformed": "1 0.2 1 2\nx 0.2 1 2\n2 0.2 3 4\n",
        "negative_first": "1 0.2 1 2\n-3 0.2 1 2\n",
        "short_line": "1 0.2 1\n2 0.27 1 2 0.3 0.1 0.2 0.3\n",
        "tabs_crlf": "7\t0.27\t1.5\t-2.5\t0.3\t1.5708\t-0\t0\r\n8 0.27 1 1 0.3 0 0 1.0\r\n",
        "full_only": "5 0.1 1 2 3 0.5 -0.25 1.25\n6 0.2 -1 -2 -3 -0.5 0.25 -1.25\n",
    }
    with tempfile.TemporaryDirectory() as td:
        for name, text in cases.items():
            p = os.path.join(td, name + ".txt")
            open(p, "w", newline="").write(text)
            m = ref.map_load(p)
            kw["case_%s_text" % name] = np.array(text)
            for k, v in m.items():
                kw["case_%s_%s" % (name, k)] = v
            print("   map case %-16s -> %d markers" % (name, len(m["id"])))
        m = ref.map_load(os.path.join(td, "does_not_exist.txt"))
        assert len(m["id"]) == 0
    save("map_txt", **kw)


if __name__ == "__main__":
    which = sys.argv[1:] or ["synth", "scene", "c5", "map"]
    if "synth" in which:
        scenario_synth()
    if "scene" in which:
        scenario_scene()
    if "c5" in which:
        scenario_c5(50, 3, 30, "slam_c5_n153", True)
    if "c5big" in which:
        scenario_c5(500, 2, 30, "slam_c5_n1503", False)
    if "map" in which:
        scenario_map()
