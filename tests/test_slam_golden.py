"""The SLAM half of the path pinned to RUNS OF THE REFERENCE ITSELF.

tests/golden/slam_*.npz and map_txt.npz were written by tools/make_golden_slam.py from oracle/_ref: the reference's
src/aruco_slam.cpp + src/map_loader.cpp compiled unmodified (stand-in Eigen / OpenCV / ROS headers) with its OpenCV
calls routed into the cv2 4.13.0 wheel.  Checked here (CPU tier):
  * oracle/orc_pose.c  orc_make_observations  vs the reference's getObservations (aruco_slam.cpp:307-376, 437-471):
    which detections survive the range / covariance gates, values, and the priority-queue pop order;
  * oracle/orc_ekf.c   predict / update       vs mu_, sigma_, aruco_id_map after every frame (:21-74, :88-263);
  * the product's make_observation (pose_core.h, host emulation) vs the same observations;
  * the whole CPU chain (oracle detector + pose + observations + EKF) on the rendered scene frames;
  * the product's map parser (host-only C-ABI entry point) and record packing vs MapLoader / toRosPose / MarkerArrays;
  * when oracle/_ref is present (authoring container; it travels to the GPU box prebuilt): the orc_* backed reference
    build reproduces the committed vectors, i.e. the goldens are not stale.
Tolerances: observations 1e-9 (cv2's Rodrigues / projectPoints vs the restatement), state 1e-9 absolute.
"""
import os

import numpy as np
import pytest

from conftest import golden

TOL = 1e-9


def _sp(O, g):
    return O.slam_params(r2c_tx=float(g["r2c_t"][0]), r2c_ty=float(g["r2c_t"][1]), marker_length=float(g["marker_length"]),
                         useful_distance_threshold=float(g["useful_distance_threshold"]))


def _pop_order(O, obs, known_ids):
    """observations in the order the reference's std::priority_queue pops them (libstdc++ heap, restated in orc_ekf.c and
    checked through the EKF below); here only: ascending landmark index, new ones first"""
    idx = [known_ids.index(o.aruco_id) if o.aruco_id in known_ids else -1 for o in obs]
    return idx


@pytest.mark.parametrize("name", ["slam_synth", "slam_scene", "slam_c5_n153"])
def test_oracle_observations_and_ekf_vs_reference_run(oracle, name):
    O = oracle
    g = golden(name)
    sp = _sp(O, g)
    e = O.Ekf(sp)
    if name.startswith("slam_c5"):
        from aruco_slam_b200.synth import c5_state
        mu0, sg0, ids0 = c5_state(int(g["n_lm"]))
        e.set_state(mu0, sg0, ids0)
    n_gated = n_new_multi = n_stationary = 0
    for f in range(int(g["n_frames"])):
        if "enc_%d" % f in g:
            for wl, wr, dt in g["enc_%d" % f]:
                e.predict(float(wl), float(wr), float(dt))
        c, ids, rv, tv = (g["det_%s_%d" % (k, f)] for k in ("corners", "ids", "rvecs", "tvecs"))
        obs = O.make_observations(c, ids, rv, tv, g["K"], g["D"], sp)
        # the same observations as the reference kept (as a multiset: its queue reorders them)
        ref_ids, ref_idx, ref_xyt, ref_cov = (g["obs_%s_%d" % (k, f)] for k in ("id", "index", "xyt", "cov"))
        assert sorted(o.aruco_id for o in obs) == sorted(ref_ids.tolist()), (name, f)
        n_gated += len(ids) - len(obs)
        for o in obs:
            j = [k for k in range(len(ref_ids)) if ref_ids[k] == o.aruco_id and abs(ref_xyt[k, 0] - o.x) < 1e-6]
            assert j, (name, f, o.aruco_id)
            assert np.abs(ref_xyt[j[0]] - [o.x, o.y, o.theta]).max() < TOL
            assert np.abs(ref_cov[j[0]] - np.array(o.cov[:])).max() < TOL
        n_new_multi += int((ref_idx < 0).sum() >= 3)
        mu_before = e.get_state()[0].copy()
        e.update(obs, dense=True)
        mu, sg, lm_ids = e.get_state()
        assert len(mu) == len(g["mu_%d" % f]), (name, f)
        if "ids_%d" % f in g:
            assert np.array_equal(lm_ids, g["ids_%d" % f]), (name, f)          # landmark order = the reference's heap order
        assert np.abs(mu - g["mu_%d" % f]).max() < TOL, (name, f)
        assert np.abs(sg - g["sigma_%d" % f]).max() < TOL, (name, f)
        if len(obs) and len(mu) == len(mu_before) and np.array_equal(mu, mu_before):
            n_stationary += 1
    if name == "slam_synth":
        assert n_gated >= 10 and n_new_multi >= 3 and n_stationary >= 1       # the vectors exercise the gates and the heap order
    if name == "slam_scene":
        assert n_stationary >= 1


def test_rank3_form_matches_reference_run(oracle):
    """Sigma - K (Gx Sigma), the form the CUDA kernels evaluate, against the reference's (I - K Gx) Sigma"""
    O = oracle
    g = golden("slam_synth")
    sp = _sp(O, g)
    e = O.Ekf(sp)
    for f in range(int(g["n_frames"])):
        for wl, wr, dt in g["enc_%d" % f]:
            e.predict(float(wl), float(wr), float(dt))
        obs = O.make_observations(*(g["det_%s_%d" % (k, f)] for k in ("corners", "ids", "rvecs", "tvecs")), g["K"], g["D"], sp)
        e.update(obs, dense=False)
    mu, sg, _ = e.get_state()
    F = int(g["n_frames"]) - 1
    assert np.abs(mu - g["mu_%d" % F]).max() < TOL and np.abs(sg - g["sigma_%d" % F]).max() < TOL


def test_c5_1503_summaries(oracle):
    """500 landmarks (state dimension 1503), 30 corrections per frame: the reference run's summaries of Sigma"""
    O = oracle
    g = golden("slam_c5_n1503")
    from aruco_slam_b200.synth import c5_state
    sp = _sp(O, g)
    e = O.Ekf(sp)
    e.set_state(*c5_state(int(g["n_lm"])))
    for f in range(int(g["n_frames"])):
        obs = O.make_observations(*(g["det_%s_%d" % (k, f)] for k in ("corners", "ids", "rvecs", "tvecs")), g["K"], g["D"], sp)
        assert len(obs) == len(g["obs_id_%d" % f]) == 30
        e.update(obs, dense=False)
        mu, sg, _ = e.get_state()
        n_lm = int(g["n_lm"])
        assert np.abs(mu - g["mu_%d" % f]).max() < TOL
        assert np.abs(np.diag(sg) - g["sigma_diag_%d" % f]).max() < TOL
        assert np.abs(sg[[0, 1, 2, 3 + 3 * (n_lm // 2), len(mu) - 1]] - g["sigma_rows_%d" % f]).max() < TOL
        assert np.abs(sg[::37, ::41] - g["sigma_sample_%d" % f]).max() < TOL
        assert abs(np.linalg.norm(sg) - float(g["sigma_fro_%d" % f])) < 1e-7


@pytest.mark.parametrize("name", ["slam_synth", "slam_scene"])
def test_product_make_observation_vs_reference_run(oracle, name):
    """pose_core.h make_observation (compiled for the host) on the golden detections"""
    from hostemu import emu
    g = golden(name)
    sp = _sp(oracle, g)
    kept = 0
    for f in range(int(g["n_frames"])):
        c, ids, rv, tv = (g["det_%s_%d" % (k, f)] for k in ("corners", "ids", "rvecs", "tvecs"))
        ref_ids, ref_xyt, ref_cov = g["obs_id_%d" % f], g["obs_xyt_%d" % f], g["obs_cov_%d" % f]
        got = []
        for k in range(len(ids)):
            ok, o = emu.observation(c[k], ids[k], rv[k], tv[k], g["K"], g["D"], sp)
            if ok:
                got.append((int(ids[k]), o))
        assert sorted(i for i, _ in got) == sorted(ref_ids.tolist()), (name, f)
        for i, o in got:
            j = [k for k in range(len(ref_ids)) if ref_ids[k] == i and abs(ref_xyt[k, 0] - o[0]) < 1e-6]
            assert j and np.abs(ref_xyt[j[0]] - o[:3]).max() < TOL and np.abs(ref_cov[j[0]] - o[3:12]).max() < TOL
            kept += 1
    assert kept > 20


def test_full_cpu_chain_on_scene_frames(oracle):
    """frames -> oracle detector -> oracle pose -> observations -> EKF, against the reference run whose detector was cv2:
    ids / corners bit-exact, poses 1e-4, the state within the tolerance the pose differences allow"""
    from aruco_slam_b200 import dictionaries as D
    O = oracle
    g = golden("slam_scene")
    sp = _sp(O, g)
    dic = D.getPredefinedDictionary(int(g["dict_id"]))
    e = O.Ekf(sp)
    frames = g["frames"]
    for f in range(int(g["n_frames"])):
        for wl, wr, dt in g["enc_%d" % f]:
            e.predict(float(wl), float(wr), float(dt))
        c, ids, _ = O.detect(frames[f], dic)
        assert np.array_equal(ids, g["det_ids_%d" % f]) and np.array_equal(c, g["det_corners_%d" % f]), f
        rv, tv = O.estimate_pose_single_markers(c, float(g["marker_length"]), g["K"], g["D"])
        if len(ids):
            assert np.abs(tv - g["det_tvecs_%d" % f]).max() < 1e-4 and np.abs(rv - g["det_rvecs_%d" % f]).max() < 1e-4
        obs = O.make_observations(c, ids, rv, tv, g["K"], g["D"], sp)
        assert sorted(o.aruco_id for o in obs) == sorted(g["obs_id_%d" % f].tolist()), f
        e.update(obs, dense=False)
        mu, sg, lm = e.get_state()
        assert np.array_equal(lm, g["ids_%d" % f])
        assert np.abs(mu - g["mu_%d" % f]).max() < 1e-4 and np.abs(sg - g["sigma_%d" % f]).max() < 1e-4, f
    # the loop localises: estimate + start pose = ground truth of the renderer
    est = g["mu_%d" % (int(g["n_frames"]) - 1)][:3] + [2.0, -0.3, 0.0]
    assert np.abs(est - g["truth"][-1]).max() < 0.05


def test_records_vs_reference_run():
    """toRosPose packing (:378-410) and the detected-map cubes (:265-281) from the golden states"""
    from aruco_slam_b200 import formats
    g = golden("slam_synth")
    F = int(g["n_frames"])
    for f in range(F):
        mu, sg, rec = g["mu_%d" % f], g["sigma_%d" % f], g["pose_%d" % f]
        pos, q, cov = formats.pose_record(mu, sg)
        assert np.array_equal(pos, rec[:3]) and np.abs(q - rec[3:7]).max() < 1e-15 and np.array_equal(cov, rec[7:])
    mu = g["mu_%d" % (F - 1)]
    cubes = formats.detected_map_records(mu, float(g["marker_length"]))
    assert [c.id for c in cubes] == g["map_id"].tolist()
    assert np.abs(np.array([[c.x, c.y, c.z] for c in cubes]) - g["map_pos"]).max() == 0
    assert np.abs(np.array([c.q for c in cubes]) - g["map_q"]).max() < 1e-15
    assert np.abs(np.array([[c.length, c.length, 0.01] for c in cubes]) - g["map_scale"]).max() == 0


def test_map_parser_vs_reference_maploader():
    """b2a_map_parse (host-only entry point of the C ABI) against MapLoader run on the same text, incl. the reference's own
    map/map.txt (7 markers)"""
    from aruco_slam_b200 import formats
    g = golden("map_txt")
    ms = formats.parse_map(str(g["ref_text"]))
    assert len(ms) == 7 and [m.id for m in ms] == g["ref_id"].tolist() == list(range(7))
    assert np.array_equal(np.array([[m.x, m.y, m.z] for m in ms]), g["ref_position"])
    assert np.abs(np.array([m.q for m in ms]) - g["ref_orientation"]).max() < 1e-15
    assert np.array_equal(np.array([[m.length, m.length, 0.01] for m in ms]), g["ref_scale"])
    for case in ("comments_blank", "malformed", "negative_first", "short_line", "tabs_crlf", "full_only"):
        text = str(g["case_%s_text" % case])
        ms = formats.parse_map(text)
        assert [m.id for m in ms] == g["case_%s_id" % case].tolist(), case
        if not ms:
            continue
        assert np.array_equal(np.array([[m.x, m.y] for m in ms]), g["case_%s_position" % case][:, :2]), case
        # lines with all eight fields: z and the orientation are defined in the reference too (on short lines it reads
        # uninitialised roll / yaw, map_loader.cpp:65-79, so only fully specified lines are compared)
        for k, (m, line) in enumerate(zip(ms, [l for l in text.replace("\r", "").split("\n") if l.strip() and not l.strip().startswith("#")
                                                and len(l.split()) >= 4])):
            if len(line.split()) == 8:
                assert m.z == g["case_%s_position" % case][k, 2]
                assert np.abs(np.array(m.q) - g["case_%s_orientation" % case][k]).max() < 1e-15, (case, k)


def test_ref_build_reproduces_goldens(oracle):
    """oracle/_ref (the reference compiled unmodified) with the orc_* hooks instead of cv2: the committed vectors are
    reproduced from the detections -- guards against stale fixtures and pins the orc_* hooks inside the reference's flow"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    ref.use_orc_hooks()
    g = golden("slam_synth")
    r = ref.RefSlam(r2c_t=tuple(g["r2c_t"]))
    r.set_camera(g["K"], g["D"])
    t = 0.0
    r.add_encoder(1.0, 1.0, t)
    img = np.zeros((8, 8), np.uint8)
    for f in range(int(g["n_frames"])):
        for wl, wr, dt in g["enc_%d" % f]:
            t += float(dt)
            r.add_encoder(float(wl), float(wr), t)
        ref.set_replay(*(g["det_%s_%d" % (k, f)] for k in ("corners", "ids", "rvecs", "tvecs")))
        r.add_image(img)
        mu, sg, ids = r.get_state()
        assert np.array_equal(ids, g["ids_%d" % f])
        # dt = difference of accumulated clock values here, the recorded dt in the golden run: rounding only
        assert np.abs(mu - g["mu_%d" % f]).max() < 1e-9 and np.abs(sg - g["sigma_%d" % f]).max() < 1e-9
    ref.set_replay()
    r.close()


def test_oracle_ekf_equals_reference_build_on_random_sequences():
    """beyond the fixed fixtures: random noise parameters, offsets, encoder messages and detections (duplicate ids in a frame,
    out-of-range and noisy markers) through oracle/_ref (the reference compiled unmodified) and through orc_pose.c / orc_ekf.c --
    the leg of tools/soak_parity.py, 12 sequences of 25 frames here"""
    import importlib.util
    import os
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    spec = importlib.util.spec_from_file_location("soak_parity", os.path.join(os.path.dirname(__file__), "..", "tools", "soak_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.soak_ekf(12) == 0
