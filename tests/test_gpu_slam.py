"""-m gpu: the SLAM half of the path (observation mapping, EKF predict / correct / augment with the
covariance resident in HBM) through the C ABI, against
  * tests/golden/slam_*.npz -- runs of the reference's own src/aruco_slam.cpp (compiled unmodified, oracle/_ref) with
    cv2 4.13.0 behind its OpenCV calls (tools/make_golden_slam.py): observations, gates, queue order, mu / Sigma per frame;
  * the CPU oracle (oracle/orc_ekf.c, itself pinned to those files in tests/test_slam_golden.py) on further scenarios.
Tolerances: FP64 throughout, same update order -> 1e-9 absolute on mu and Sigma (north_star asks 1e-4)."""
import numpy as np
import pytest

from conftest import golden
from test_oracle_ekf import _scenario
from aruco_slam_b200 import dictionaries as D, synth

pytestmark = pytest.mark.gpu


def _obs_struct(cls, aid, x, y, th, R):
    o = cls()
    o.aruco_id, o.aruco_index, o.x, o.y, o.theta = int(aid), -1, float(x), float(y), float(th)
    for i in range(9):
        o.cov[i] = float(np.asarray(R).reshape(9)[i])
    return o


@pytest.mark.parametrize("seed", [0, 1])
def test_ekf_scenario_vs_oracle(oracle, seed):
    """predict + correct + augment over 25 frames, duplicates and repeated observations included"""
    from aruco_slam_b200 import slam, _lib
    s = slam.ArucoSlam(image_shape=(64, 64))
    e = oracle.Ekf(oracle.slam_params())
    s.addEncoder(0, 0, None)                         # the first message only latches (aruco_slam.cpp:24-29)
    for (wl, wr, dt), obs in _scenario(seed):
        s.addEncoder(wl, wr, dt)
        e.predict(wl, wr, dt)
        s.update([_obs_struct(_lib.Observation, *o) for o in obs])
        e.update([_obs_struct(oracle.Observation, *o) for o in obs])
        mu, sg, ids = s.get_state()
        omu, osg, oids = e.get_state()
        assert np.array_equal(ids, oids)
        assert np.abs(mu - omu).max() < 1e-9
        assert np.abs(sg - osg).max() < 1e-9
    assert len(mu) > 3 + 3 * 8
    s.close()


@pytest.mark.parametrize("n_lm", [334, 500])          # state dimension 1005 and 1503 (BASELINE config 5, both readings)
def test_ekf_large_state_vs_oracle(oracle, n_lm):
    """30 observations of distinct known landmarks on a 1003+ / 1503-dimensional filter (set_state / get_state round trip)"""
    from aruco_slam_b200 import slam, _lib
    rng = np.random.default_rng(5)
    N = 3 + 3 * n_lm
    A = rng.normal(size=(N, N))
    sigma0 = A @ A.T / N + 0.1 * np.eye(N)
    mu0 = np.concatenate([[0.3, -0.2, 0.4], rng.uniform(-4, 4, 3 * n_lm)])
    ids = np.arange(n_lm, dtype=np.int32) * 2 + 1
    s = slam.ArucoSlam(image_shape=(64, 64), max_landmarks=n_lm + 4)
    e = oracle.Ekf(oracle.slam_params())
    s.set_state(mu0, sigma0, ids)
    e.set_state(mu0, sigma0, ids)
    mu, sg, gi = s.get_state()
    assert np.array_equal(mu, mu0) and np.array_equal(sg, sigma0) and np.array_equal(gi, ids)
    obs = []
    for k in rng.choice(n_lm, 30, replace=False):
        L = 3 + 3 * k
        c, sn = np.cos(mu0[2]), np.sin(mu0[2])
        dx, dy = mu0[L] - mu0[0], mu0[L + 1] - mu0[1]
        z = np.array([dx * c + dy * sn, -dx * sn + dy * c, mu0[L + 2] - mu0[2]]) + rng.normal(0, 0.02, 3)
        obs.append((ids[k], z[0], z[1], z[2], np.diag([0.02, 0.02, 0.003])))
    obs.append((9999, 1.0, 0.5, 0.2, np.diag([0.02, 0.02, 0.003])))          # and one new landmark
    s.update([_obs_struct(_lib.Observation, *o) for o in obs])
    e.update([_obs_struct(oracle.Observation, *o) for o in obs])
    mu, sg, gi = s.get_state()
    omu, osg, oi = e.get_state()
    assert len(mu) == N + 3 and np.array_equal(gi, oi)
    assert np.abs(mu - omu).max() < 1e-9
    assert np.abs(sg - osg).max() < 1e-9
    s.close()


def test_add_image_vs_oracle_pipeline(oracle):
    """addImage(img): detect + pose + observation mapping + EKF on the GPU == the same chain through the oracle"""
    from aruco_slam_b200 import slam
    dic_id = D.DICT_4X4_50
    K = np.array([[600.0, 0, 320], [0, 600.0, 240], [0, 0, 1]])
    Dist = np.array([0.05, -0.1, 0.001, -0.002, 0.02])
    L = 0.27
    # four markers at true 3-D poses (planar squares), so the reprojection error and hence the observation covariance are small
    # (a marker facing the camera has a rotation of about pi around x: object y points up, image y down)
    poses = [(3, [3.0, -0.3, 0.05], [-0.6, -0.3, 2.2]), (11, [2.8, 0.4, 0.3], [0.7, -0.2, 2.6]),
             (24, [3.05, 0.2, -0.4], [-0.5, 0.45, 2.0]), (40, [2.9, 0.1, 0.2], [0.6, 0.4, 2.4])]
    fr = synth.render_scene(640, 480, dic_id, K, Dist, L, poses).image
    s = slam.ArucoSlam(dic_id, L, image_shape=fr.shape, useful_distance_threshold=50.0)
    s.setCameraParameters(K, Dist)
    s.addImage(fr)                                   # before the first encoder message: ignored (aruco_slam.cpp:84-85)
    assert s.dim == 3
    s.addEncoder(0, 0, None)
    sp = oracle.slam_params(marker_length=L, useful_distance_threshold=50.0)
    e = oracle.Ekf(sp)
    dic = D.getPredefinedDictionary(dic_id)
    for step in range(3):
        s.addEncoder(2.0, 2.5, 0.1)
        e.predict(2.0, 2.5, 0.1)
        s.addImage(fr)
        oc, oi, _ = oracle.detect(fr, dic)
        orv, otv = oracle.estimate_pose_single_markers(oc, L, K, Dist)
        obs = oracle.make_observations(oc, oi, orv, otv, K, Dist, sp)
        e.update(obs)
        mu, sg, ids = s.get_state()
        omu, osg, oids = e.get_state()
        assert len(oids) == 4 and np.array_equal(ids, oids)
        assert np.abs(mu - omu).max() < 1e-4        # north_star: 1e-4 m / rad (poses come from two LM implementations)
        assert np.abs(sg - osg).max() < 1e-4
    s.close()


def test_pose_and_map_records():
    """the ROS-free records of the filter state: toRosPose's packing (aruco_slam.cpp:376-407) and the detected-map cubes (:266-281)"""
    from scipy.spatial.transform import Rotation
    from aruco_slam_b200 import slam, formats
    rng = np.random.default_rng(5)
    n_lm = 4
    N = 3 + 3 * n_lm
    A = rng.normal(size=(N, N))
    sigma = A @ A.T / N + 0.1 * np.eye(N)
    mu = rng.uniform(-2, 2, N)
    s = slam.ArucoSlam(marker_length=0.27, image_shape=(64, 64))
    s.set_state(mu, sigma, np.arange(10, 10 + n_lm, dtype=np.int32))
    p = formats.robot_pose(s)
    assert np.array_equal(p.position, [mu[0], mu[1], 0.1])
    q = Rotation.from_euler("xyz", [0, 0, mu[2]]).as_quat()
    assert np.allclose(p.orientation, q if np.dot(q, p.orientation) >= 0 else -q, atol=1e-14)
    want = np.zeros((6, 6))
    want[np.ix_([0, 1, 5], [0, 1, 5])] = sigma[:3, :3]
    assert np.array_equal(p.covariance, want)
    cubes = formats.detected_map(s)
    assert [c.id for c in cubes] == list(range(n_lm))             # the landmark index, as the reference publishes it
    for i, c in enumerate(cubes):
        assert (c.length, c.x, c.y, c.z) == (0.27, mu[3 + 3 * i], mu[4 + 3 * i], 0.3)
        q = Rotation.from_euler("xyz", [0, 1.5708, mu[5 + 3 * i]]).as_quat()
        assert np.allclose(c.q, q if np.dot(q, c.q) >= 0 else -q, atol=1e-14)
    s.close()


def _golden_slam(g, **kw):
    from aruco_slam_b200 import slam
    s = slam.ArucoSlam(int(g["dict_id"]) if "dict_id" in g else 16, float(g["marker_length"]), r2c_tx=float(g["r2c_t"][0]), r2c_ty=float(g["r2c_t"][1]),
                       useful_distance_threshold=float(g["useful_distance_threshold"]), **kw)
    s.setCameraParameters(g["K"], g["D"])
    return s


COV_TOL = 1e-7    # CalculateCovariance rounds the projected corners to float (cv::Point2f, aruco_slam.cpp:440-442): a 1e-16
                  # difference in the double projection (FMA contraction on the GPU) can flip that rounding by one float ulp
                  # (3e-5 px), which moves the reprojection-error term of the covariance by up to ~1e-8


def _check_obs(obs, g, f):
    """the kept observations == the reference's (its queue reorders them, so compare as a set keyed by id and x); returns
    the reference's own values in detection order (what getObservations pushed), as C structs"""
    from aruco_slam_b200 import _lib
    ref_ids, ref_xyt, ref_cov = g["obs_id_%d" % f], g["obs_xyt_%d" % f], g["obs_cov_%d" % f]
    assert sorted(o.aruco_id for o in obs) == sorted(ref_ids.tolist()), f
    out = []
    for o in obs:
        j = [k for k in range(len(ref_ids)) if ref_ids[k] == o.aruco_id and abs(ref_xyt[k, 0] - o.x) < 1e-6]
        assert j, (f, o.aruco_id)
        assert np.abs(ref_xyt[j[0]] - [o.x, o.y, o.theta]).max() < 1e-9
        assert np.abs(ref_cov[j[0]] - np.array(o.cov[:])).max() < COV_TOL
        out.append(_obs_struct(_lib.Observation, o.aruco_id, *ref_xyt[j[0]], ref_cov[j[0]]))
    return out


def test_observations_and_ekf_vs_reference_run_synth():
    """slam_synth: detections replayed into the reference -- range gate and covariance gate rejecting (aruco_slam.cpp:327-333,
    367-368), >= 3 new landmarks in one frame (heap order), duplicates, repeated frames (stationary gate :192-198)"""
    g = golden("slam_synth")
    s = _golden_slam(g, image_shape=(64, 64))
    s.update([])                                       # nothing before the first encoder message changes the state
    s.addEncoder(1.0, 1.0, 0.05)                       # latch only (:24-29)
    mu, sg, _ = s.get_state()
    assert len(mu) == 3 and not mu.any() and not sg.any()
    n_gated = 0
    for f in range(int(g["n_frames"])):
        for wl, wr, dt in g["enc_%d" % f]:
            s.addEncoder(wl, wr, dt)
        c, ids, rv, tv = (g["det_%s_%d" % (k, f)] for k in ("corners", "ids", "rvecs", "tvecs"))
        obs = s.make_observations(c, ids, rv, tv)
        n_gated += len(ids) - len(obs)
        s.update(_check_obs(obs, g, f))                        # the filter is fed the reference's own observation values
        mu, sg, lm = s.get_state()
        assert np.array_equal(lm, g["ids_%d" % f]), f          # landmark order = the reference's priority-queue order
        assert np.abs(mu - g["mu_%d" % f]).max() < 1e-9 and np.abs(sg - g["sigma_%d" % f]).max() < 1e-9, f
    assert n_gated >= 10
    from aruco_slam_b200 import formats
    p = formats.robot_pose(s)
    rec = g["pose_%d" % (int(g["n_frames"]) - 1)]
    assert np.abs(np.concatenate([p.position, p.orientation, p.covariance.ravel()]) - rec).max() < 1e-9
    cubes = formats.detected_map(s)
    assert [c.id for c in cubes] == g["map_id"].tolist()
    assert np.abs(np.array([[c.x, c.y, c.z] for c in cubes]) - g["map_pos"]).max() < 1e-9
    assert np.abs(np.array([c.q for c in cubes]) - g["map_q"]).max() < 1e-9
    s.close()


def test_add_image_vs_reference_run_scene():
    """slam_scene: the reference's addImage on rendered frames of a map.txt-style room (DICT_ARUCO_ORIGINAL, the intrinsics
    of /root/reference/default.yaml:10-20), cv2 behind its detectMarkers / solvePnP.  addImage on the GPU: ids / corners
    bit-exact, poses 1e-4, and the filter state within what the pose tolerance allows (1e-4)."""
    g = golden("slam_scene")
    frames = g["frames"]
    s = _golden_slam(g, image_shape=frames.shape[1:])
    s.addEncoder(0, 0, None)
    for f in range(int(g["n_frames"])):
        for wl, wr, dt in g["enc_%d" % f]:
            s.addEncoder(wl, wr, dt)
        r = s.detector.detect_pose_batch(frames[f], float(g["marker_length"]), g["K"], g["D"])
        assert np.array_equal(r.ids[0], g["det_ids_%d" % f]) and np.array_equal(r.corners[0], g["det_corners_%d" % f]), f
        if len(r.ids[0]):
            assert np.abs(r.tvecs[0] - g["det_tvecs_%d" % f]).max() < 1e-4
            assert max(synth.rvec_distance(a, b) for a, b in zip(r.rvecs[0], g["det_rvecs_%d" % f])) < 1e-4
        obs = s.make_observations(r.corners[0], r.ids[0], r.rvecs[0], r.tvecs[0])
        assert sorted(o.aruco_id for o in obs) == sorted(g["obs_id_%d" % f].tolist()), f
        s.addImage(frames[f])
        mu, sg, lm = s.get_state()
        assert np.array_equal(lm, g["ids_%d" % f]), f
        assert np.abs(mu - g["mu_%d" % f]).max() < 1e-4 and np.abs(sg - g["sigma_%d" % f]).max() < 1e-4, f
        if f in (0, 7):                              # getMarkedImg (aruco_slam.h:152): the frame with the detections drawn on it
            marked = s.getMarkedImg()
            want = s.detector.drawDetectedMarkers(frames[f].copy(), g["det_corners_%d" % f], g["det_ids_%d" % f])
            assert np.array_equal(marked, want) and not np.array_equal(marked, frames[f])
    s.close()


def test_add_image_on_golden_detections_is_exact():
    """the same scene with the reference run's own detections through make_observations (xyt 1e-9, covariance 1e-7) and the
    reference's observation values through update: mu, Sigma 1e-9"""
    g = golden("slam_scene")
    s = _golden_slam(g, image_shape=(64, 64))
    s.addEncoder(0, 0, None)
    for f in range(int(g["n_frames"])):
        for wl, wr, dt in g["enc_%d" % f]:
            s.addEncoder(wl, wr, dt)
        obs = s.make_observations(*(g["det_%s_%d" % (k, f)] for k in ("corners", "ids", "rvecs", "tvecs")))
        s.update(_check_obs(obs, g, f))
        mu, sg, lm = s.get_state()
        assert np.array_equal(lm, g["ids_%d" % f])
        assert np.abs(mu - g["mu_%d" % f]).max() < 1e-9 and np.abs(sg - g["sigma_%d" % f]).max() < 1e-9, f
    s.close()


@pytest.mark.parametrize("name", ["slam_c5_n153", "slam_c5_n1503"])
def test_c5_vs_reference_run(name):
    """EKF-only workload (BASELINE config 5): 30 corrections of known landmarks per frame from a dense SPD Sigma0"""
    g = golden(name)
    n_lm = int(g["n_lm"])
    s = _golden_slam(g, image_shape=(64, 64), max_landmarks=n_lm + 4)
    s.set_state(*synth.c5_state(n_lm))
    for f in range(int(g["n_frames"])):
        obs = s.make_observations(*(g["det_%s_%d" % (k, f)] for k in ("corners", "ids", "rvecs", "tvecs")))
        assert len(obs) == 30
        s.update(_check_obs(obs, g, f))
        mu, sg, _ = s.get_state()
        assert np.abs(mu - g["mu_%d" % f]).max() < 1e-9
        if "sigma_%d" % f in g:
            assert np.abs(sg - g["sigma_%d" % f]).max() < 1e-9
        else:
            assert np.abs(np.diag(sg) - g["sigma_diag_%d" % f]).max() < 1e-9
            assert np.abs(sg[[0, 1, 2, 3 + 3 * (n_lm // 2), len(mu) - 1]] - g["sigma_rows_%d" % f]).max() < 1e-9
            assert np.abs(sg[::37, ::41] - g["sigma_sample_%d" % f]).max() < 1e-9
            assert abs(np.linalg.norm(sg) - float(g["sigma_fro_%d" % f])) < 1e-7
    s.close()


@pytest.mark.parametrize("depth", [2, 4, 8])
def test_add_image_submit_wait_equals_sequential_loop(depth):
    """b2a_slam_add_image_submit / _wait with 2 / 4 / 8 frames in flight: the same filter state, frame by frame, as the sequential addImage loop"""
    import collections
    from aruco_slam_b200.aruco import ArucoDetector
    g = golden("slam_scene")
    frames = g["frames"]
    n = int(g["n_frames"])
    a = _golden_slam(g, image_shape=frames.shape[1:])
    b = _golden_slam(g, image_shape=frames.shape[1:])
    descs = [ArucoDetector._frames_host(frames[f]) for f in range(n)]
    b.detector.set_inflight(depth)
    assert b.submitImageFrames(descs[0][0]) == -1                  # before the first encoder message the frame is ignored (aruco_slam.cpp:84-85)
    b.waitImage(-1)
    a.addEncoder(0, 0, None); b.addEncoder(0, 0, None)
    q, nxt = collections.deque(), 0
    for f in range(n):
        while nxt < n and len(q) < depth:
            q.append(b.submitImageFrames(descs[nxt][0]))
            nxt += 1
        for wl, wr, dt in g["enc_%d" % f]:
            a.addEncoder(wl, wr, dt); b.addEncoder(wl, wr, dt)
        a.addImage(frames[f])
        b.waitImage(q.popleft())
        ma, sa, ia = a.get_state()
        mb, sb, ib = b.get_state()
        assert np.array_equal(ia, ib) and np.array_equal(ma, mb) and np.array_equal(sa, sb), f
        assert np.array_equal(ia, g["ids_%d" % f])
    a.close(); b.close()


def test_robot_pose_submit_wait_equals_synchronous_read():
    """b2a_slam_robot_pose_submit / _wait: the record enqueued behind a frame's filter work equals the synchronous read at that point"""
    from aruco_slam_b200 import formats
    g = golden("slam_scene")
    frames = g["frames"]
    s = _golden_slam(g, image_shape=frames.shape[1:])
    s.addEncoder(0, 0, None)
    want = []
    for f in range(6):
        for wl, wr, dt in g["enc_%d" % f]:
            s.addEncoder(wl, wr, dt)
        s.addImage(frames[f])
        formats.robot_pose_submit(s, f % 8)
        want.append(formats.robot_pose(s))
    for f in range(6):
        got = formats.robot_pose_wait(s, f % 8)
        assert np.array_equal(got.position, want[f].position) and np.array_equal(got.orientation, want[f].orientation)
        assert np.array_equal(got.covariance, want[f].covariance)
    with pytest.raises(Exception):
        formats.robot_pose_wait(s, 7)                                   # nothing submitted in that slot
    s.close()
