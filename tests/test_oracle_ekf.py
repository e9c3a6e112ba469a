"""The C oracle EKF vs an independent NumPy restatement in the reference's own
matrix form.  PARITY UNPINNED by the reference (no tests/golden there): this
only shows two independent readings of src/aruco_slam.cpp:21-74,88-263 agree."""
import numpy as np
import pytest

from ekf_numpy import NumpyEkf


def _scenario(seed, n_lm=12, frames=25):
    rng = np.random.default_rng(seed)
    lms = {int(i): (rng.uniform(-4, 4), rng.uniform(-4, 4), rng.uniform(-3, 3)) for i in rng.choice(200, n_lm, replace=False)}
    pose = np.zeros(3)
    for f in range(frames):
        wl, wr, dt = rng.uniform(0, 6), rng.uniform(0, 6), 0.1
        ds = 0.05 * dt * (wl + wr) / 2
        dth = 0.05 * dt * (wr - wl) / 0.18
        pose = pose + [ds * np.cos(pose[2] + dth / 2), ds * np.sin(pose[2] + dth / 2), dth]
        obs = []
        for aid, (mx, my, mth) in lms.items():
            if rng.random() < 0.5:
                continue
            c, s = np.cos(pose[2]), np.sin(pose[2])
            dx, dy = mx - pose[0], my - pose[1]
            z = np.array([dx * c + dy * s, -dx * s + dy * c, mth - pose[2]]) + rng.normal(0, 0.01, 3)
            oe = rng.uniform(0, 2e-4)
            R = np.diag([oe * 100 + 1e-2, oe * 100 + 1e-2, oe * 10 + 1e-3])
            obs.append((aid, z[0], z[1], z[2], R))
        if f % 7 == 3 and obs:
            obs.append(obs[0])          # duplicate detection in one frame
        yield (wl, wr, dt), obs


@pytest.mark.parametrize("dense", [True, False])
@pytest.mark.parametrize("seed", [0, 1])
def test_ekf_oracle_vs_numpy(oracle, seed, dense):
    sp = oracle.slam_params()
    e = oracle.Ekf(sp)
    ref = NumpyEkf()
    for (wl, wr, dt), obs in _scenario(seed):
        e.predict(wl, wr, dt)
        ref.predict(wl, wr, dt)
        arr = []
        for aid, x, y, th, R in obs:
            o = oracle.Observation()
            o.aruco_id, o.aruco_index, o.x, o.y, o.theta = aid, -1, x, y, th
            for i in range(9):
                o.cov[i] = R.reshape(9)[i]
            arr.append(o)
        e.update(arr, dense=dense)
        ref.update(obs)
        mu, sg, ids = e.get_state()
        assert len(mu) == len(ref.mu)
        assert [ref.id_map[i] for i in ids.tolist()] == [ids.tolist().index(i) for i in ids.tolist()]
        assert np.abs(mu - ref.mu).max() < 1e-9
        assert np.abs(sg - ref.sigma).max() < 1e-9
    assert len(mu) > 3 + 3 * 8


def test_repeated_observation_is_gated(oracle):
    """aruco_slam.cpp:192-198: the same observation twice in consecutive frames skips the update."""
    sp = oracle.slam_params()
    e = oracle.Ekf(sp)
    e.predict(1, 1, 0.1)      # latch (no-op on zero state besides motion)
    def ob(aid, x, y, th):
        o = oracle.Observation(); o.aruco_id = aid; o.aruco_index = -1; o.x = x; o.y = y; o.theta = th
        for i, v in enumerate([0.02, 0, 0, 0, 0.02, 0, 0, 0, 0.003]): o.cov[i] = v
        return o
    e.update([ob(3, 1.0, 0.2, 0.1)])           # new landmark (last_observation_ unset)
    e.predict(2, 1, 0.1)
    e.update([ob(3, 1.01, 0.2, 0.1)])          # full update, records z
    mu1, sg1, _ = e.get_state()
    e.update([ob(3, 1.01, 0.2, 0.1)])          # identical -> gated, state untouched
    mu2, sg2, _ = e.get_state()
    assert np.array_equal(mu1, mu2) and np.array_equal(sg1, sg2)
    e.update([ob(3, 1.01, 0.2, 0.1)])          # previous frame left it unset -> updates again
    mu3, _, _ = e.get_state()
    assert not np.array_equal(mu2, mu3)
