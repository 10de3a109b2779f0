"""CPU: the oracle IPM against (a) the committed golden fixtures, (b) an independent KKT certificate built
from torch.autograd derivatives of the literal NLP, (c) the known answer of SURVEY App. D."""
from pathlib import Path

import numpy as np
import pytest

from oracle import nlp_ref

GOLD = Path(__file__).resolve().parent / "golden"
NAMES = ["nmpc_tt", "t_trajectory", "plus_trajectory", "race_trajectory_1", "race_track_2", "10_obstacles"]


def _setup(pkg, oracle_mod, name):
    sc = pkg.SCENARIOS[name]
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    return sc, sp, sc.obstacle_table(), sc.bounds()


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(pkg, oracle_mod, name):
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, name)
    G = np.load(GOLD / f"solves_{name}.npz")
    r = oracle_mod.solve(sp, obs, G["p"], G["x0"], lbx, ubx, lbg, ubg)
    assert np.array_equal(r["status"], G["status"])
    assert np.array_equal(r["iters"], G["iters"])
    ok = G["status"] == 0
    assert np.allclose(r["f"][ok], G["f"][ok], rtol=1e-12, atol=0)
    assert np.allclose(r["x"][ok], G["x"][ok], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", ["t_trajectory", "race_track_2", "nmpc_tt"])
def test_kkt_certificate(pkg, oracle_mod, name):
    """Converged golden solutions satisfy the KKT conditions of the literal NLP (derivatives by autograd)."""
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, name)
    rs = nlp_ref.RefSpec(T=sc.T, N=sc.N, obstacles=sc.obstacles, uav_r=sc.uav_r, w1=sc.w1, w2=sc.w2)
    G = np.load(GOLD / f"solves_{name}.npz")
    idx = [i for i in range(len(G["status"])) if G["status"][i] == 0][:3]
    assert idx
    for i in idx:
        x, lam_g, lam_x = G["x"][i], G["lam_g"][i], G["lam_x"][i]
        d = nlp_ref.eval_all(rs, x, G["p"][i])
        assert abs(d["f"] - G["f"][i]) <= 1e-10 * abs(d["f"])
        assert np.abs(d["g"] - G["g"][i]).max() <= 1e-9
        stat = d["grad"] + d["J"].T @ lam_g + lam_x
        # IPOPT's test is on the SCALED error: tol * s_d with s_d = max(100, mean |multiplier|) / 100,
        # divided by the objective scaling; the unscaled absolute tolerance (dual_inf_tol) is 1.
        mult = max(1.0, np.abs(lam_g).max(), np.abs(lam_x).max())
        assert np.abs(stat).max() <= 1e-6 + 1e-7 * mult, (np.abs(stat).max(), mult)
        # primal feasibility (IPOPT: constr_viol_tol 1e-4 absolute, tol 1e-8 scaled, bounds relaxed by 1e-8)
        assert np.all(d["g"] <= ubg + 1e-6) and np.all(d["g"] >= lbg - 1e-6)
        assert np.all(x <= ubx + 1e-12) and np.all(x >= lbx - 1e-12)
        # sign + complementarity: lam > 0 only on active upper bounds, < 0 only on active lower bounds
        su, sl = ubg - d["g"], d["g"] - lbg
        assert np.all(lam_g[su > 1e-3] <= 1e-5) and np.all(lam_g[sl > 1e-3] >= -1e-5)
        xu, xl = ubx - x, x - lbx
        assert np.all(lam_x[xu > 1e-4] <= 1e-4) and np.all(lam_x[xl > 1e-4] >= -1e-4)


def test_known_answer_first_solve_T02(pkg, oracle_mod):
    """the first NLP of the T=0.2 scripts (x0=[99,150,80,0..], target (100,150)) has the mirrored
    minima f* ~= 248.10109 with u0* ~= [14, +pi/30, +-pi/21, ~0, -pi/30, ~0]."""
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, "t_trajectory")
    p = np.array(list(sc.x_init) + list(sc.target_init))
    r = oracle_mod.solve(sp, obs, p, np.zeros(sc.n_w), lbx, ubx, lbg, ubg, nthreads=1)
    assert r["status"][0] == 0
    assert abs(r["f"][0] - 248.10109) < 2e-5
    u0 = r["x"][0][:6]
    assert abs(u0[0] - 14) < 1e-6 and abs(u0[1] - np.pi / 30) < 1e-6 and abs(abs(u0[2]) - np.pi / 21) < 1e-6
    assert abs(u0[4] + np.pi / 30) < 1e-6


def test_empty_and_single(pkg, oracle_mod):
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, "t_trajectory")
    r = oracle_mod.solve(sp, obs, np.zeros((0, 11)), np.zeros((0, sc.n_w)), lbx, ubx, lbg, ubg)
    assert r["x"].shape == (0, sc.n_w)
