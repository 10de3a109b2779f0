"""CPU: the C++ oracle's f, g, grad f, J, Hess_L against torch.autograd on the literal restatement
(oracle/nlp_ref.py) of Python/NMPC_TT.py:139-148, :160-167, :193-221, :234-244."""
import numpy as np
import pytest

from oracle import nlp_ref

RT2 = [(0, 80), (500, 245), (1000, 70), (1500, 295), (1765, 550), (1500, 750), (1000, 1005), (500, 800), (-100, 950), (-200, 550)]
SPECS = {
    "nmpc_tt": nlp_ref.RefSpec(T=1.0),
    "t02": nlp_ref.RefSpec(T=0.2, obstacles=((10000.0, 10000.0, 30.0),) * 3),
    "rt2": nlp_ref.RefSpec(T=0.2, obstacles=tuple((float(a), float(b), 50.0) for a, b in RT2)),
    "short": nlp_ref.RefSpec(T=0.5, N=4, obstacles=((120.0, 160.0, 10.0),), w1=1.5, w2=3.0),
}


def _point(rs, seed):
    rng = np.random.default_rng(seed)
    lbx, ubx, _, _ = nlp_ref.bounds(rs)
    w = lbx + rng.random(rs.n_w) * (ubx - lbx)
    p = np.array([90 + rng.uniform(-30, 30), 150 + rng.uniform(-30, 30), rng.uniform(80, 140), rng.uniform(-.2, .2),
                  rng.uniform(-3, 3), rng.uniform(-.4, .4), rng.uniform(-.4, .4), rng.uniform(-1, 1),
                  100 + rng.uniform(-40, 40), 150 + rng.uniform(-40, 40), rng.uniform(-3, 3)])
    lam = rng.standard_normal(rs.n_g)
    return w, p, lam


@pytest.mark.parametrize("name", list(SPECS))
def test_oracle_matches_autograd(oracle_mod, name):
    rs = SPECS[name]
    sp = oracle_mod.make_spec(rs.T, rs.N, rs.n_obs, rs.w1, rs.w2, rs.vfov, rs.hfov)
    obs = oracle_mod.obstacle_table(rs.obstacles, rs.uav_r)
    w, p, lam = _point(rs, 7)
    ref = nlp_ref.eval_all(rs, w, p, lam, 0.6)
    got = oracle_mod.evaluate(sp, obs, w, p, lam, 0.6, hessian=True)
    for key, tol in [("f", 1e-13), ("g", 1e-13), ("grad", 1e-12), ("J", 1e-12), ("H", 1e-11)]:
        a, b = np.asarray(ref[key]), np.asarray(got[key])
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(a).max()), key


@pytest.mark.parametrize("name", ["nmpc_tt", "short"])
def test_oracle_matches_autograd_with_target_trajectory(oracle_mod, name):
    """per-stage predicted target instead of the constant (p[8], p[9])."""
    rs = SPECS[name]
    sp = oracle_mod.make_spec(rs.T, rs.N, rs.n_obs, rs.w1, rs.w2, rs.vfov, rs.hfov)
    obs = oracle_mod.obstacle_table(rs.obstacles, rs.uav_r)
    w, p, lam = _point(rs, 19)
    k = np.arange(rs.N)[:, None]
    traj = np.array([p[8], p[9]]) + rs.T * k * np.array([[9.0, -4.0]]) + 0.3 * np.sin(k)        # a curved prediction
    ref = nlp_ref.eval_all(rs, w, p, lam, 0.6, target_traj=traj)
    got = oracle_mod.evaluate(sp, obs, w, p, lam, 0.6, hessian=True, target_traj=traj)
    for key, tol in [("f", 1e-13), ("g", 1e-13), ("grad", 1e-12), ("J", 1e-12), ("H", 1e-11)]:
        a, b = np.asarray(ref[key]), np.asarray(got[key])
        assert np.abs(a - b).max() <= tol * max(1.0, np.abs(a).max()), key
    # a trajectory that repeats p[8:10] is the reference problem
    same = oracle_mod.evaluate(sp, obs, w, p, lam, 0.6, hessian=True, target_traj=np.tile(p[8:10], (rs.N, 1)))
    base = oracle_mod.evaluate(sp, obs, w, p, lam, 0.6, hessian=True)
    assert same["f"] == base["f"] and np.array_equal(same["H"], base["H"])


def test_gradient_against_central_differences(oracle_mod):
    rs = SPECS["t02"]
    sp = oracle_mod.make_spec(rs.T, rs.N, rs.n_obs)
    obs = oracle_mod.obstacle_table(rs.obstacles)
    w, p, lam = _point(rs, 3)
    got = oracle_mod.evaluate(sp, obs, w, p)
    h = 1e-6
    for i in [0, 1, 2, 7, 44, 83]:
        e = np.zeros_like(w); e[i] = h
        fd = (oracle_mod.evaluate(sp, obs, w + e, p)["f"] - oracle_mod.evaluate(sp, obs, w - e, p)["f"]) / (2 * h)
        assert abs(fd - got["grad"][i]) <= 1e-6 * max(1.0, abs(fd))


def test_structure_facts(oracle_mod):
    """last-stage controls never enter the cost; stage-0 rows of g have zero Jacobian;
    the angle rows are linear in w."""
    rs = SPECS["nmpc_tt"]
    sp = oracle_mod.make_spec(rs.T, rs.N, rs.n_obs)
    obs = oracle_mod.obstacle_table(rs.obstacles)
    w, p, lam = _point(rs, 11)
    got = oracle_mod.evaluate(sp, obs, w, p, hessian=False)
    assert np.all(got["grad"][84:] == 0.0)           # u_{N-1}
    assert np.all(got["grad"][[79, 80]] == 0.0)      # pitch / yaw rate of u_{N-2}
    assert np.all(got["J"][:8] == 0.0)               # stage-0 rows
    w2, _, _ = _point(rs, 12)
    got2 = oracle_mod.evaluate(sp, obs, w2, p)
    rows = 5 + rs.n_obs
    lin = [k * rows + i for k in range(rs.N + 1) for i in (1, 2, 3, 4)]
    assert np.allclose(got["J"][lin], got2["J"][lin], atol=0, rtol=0)


def test_bounds_layout():
    rs = SPECS["rt2"]
    lbx, ubx, lbg, ubg = nlp_ref.bounds(rs)
    assert lbx.shape == (90,) and lbg.shape == (240,)
    assert lbx[0] == 14 and ubx[6] == 30 and np.isclose(ubx[2], np.pi / 21)     # NMPC_TT.py:294-306
    assert lbg[0] == 75 and ubg[15] == 150 and lbg[5] == -np.inf and ubg[14] == 0  # Race Track 2.py:295-326
