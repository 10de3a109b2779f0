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


def casadi_fixture(name):
    """tests/golden/casadi_<name>.npz -- records of the REAL reference (CasADi + IPOPT) made by bench/run_casadi.py --
    or None.  None in this repository: CasADi cannot be installed in the build image, parity is unpinned."""
    f = GOLD / f"casadi_{name}.npz"
    return np.load(f, allow_pickle=False) if f.exists() else None


@pytest.mark.parametrize("name", NAMES)
def test_oracle_against_casadi_records(pkg, oracle_mod, name):
    """PINNING HOOK: when real-reference records exist, the oracle must reproduce IPOPT's outcome on the same (p, x0):
    identical return status, f* to 1e-8, first input u0* to 1e-6 (the north star's tolerances), and iteration counts
    within +-2 on 90 % of the solves."""
    C = casadi_fixture(name)
    if C is None:
        pytest.skip("no tests/golden/casadi_*.npz (CasADi unavailable here; run bench/run_casadi.py where it is installed)")
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, name)
    r = oracle_mod.solve(sp, obs, C["p"], C["x0"], lbx, ubx, lbg, ubg)
    assert np.array_equal(r["status"], C["status"]), (r["status"], C["status"])
    ok = C["status"] == 0
    # (a uniform f* mismatch of 1 ... 2e-8 would point at the evaluation point of sol['f'] -- projected vs unprojected final
    #  iterate, DESIGN.md section 5 "known candidate deviation" -- not at the iterates)
    assert np.all(np.abs(r["f"][ok] - C["f"][ok]) <= 1e-8 * np.abs(C["f"][ok]))
    assert np.all(np.abs(r["x"][ok, :6] - C["x"][ok, :6]).max(axis=1) <= 1e-6 * np.abs(C["x"][ok, :6]).max(axis=1))
    assert (np.abs(r["iters"][ok] - C["iters"][ok]) <= 2).mean() >= 0.9


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


# ----------------------------------------------------------------------------------------------------------------
# restoration phase, watchdog, soft restoration (IPOPT's globalisation beyond the plain filter line search)
# ----------------------------------------------------------------------------------------------------------------
def _infeasible_batch(pkg, sc, B, seed=99):
    p, _ = pkg.random_instances(sc, B, seed=seed)
    ob = sc.obstacle_table()
    h = B // 2
    p[:h, 0] = ob[1, 0] + 10.0; p[:h, 1] = ob[1, 1] - 5.0               # inside obstacle 2's keep-out disc (stage-0 row violated)
    p[h:, 2] = 150.0 + 10.0 ** np.linspace(-5.5, -2, B - h)             # above the altitude ceiling by more than the relaxation
    return p


def test_restoration_on_infeasible_nlp(pkg, oracle_mod):
    """An NLP whose stage-0 rows (functions of p only) are violated is infeasible whatever w: IPOPT enters the
    restoration phase and ends with Infeasible_Problem_Detected (6), Restoration_Failed (2) or max_iter (1).  With the
    restoration phase switched off the same instances end with Restoration_Failed at the first failed line search."""
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, "nmpc_tt")
    B = 16
    p = _infeasible_batch(pkg, sc, B)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    r = oracle_mod.solve(sp, obs, p, x0, lbx, ubx, lbg, ubg)
    assert np.isin(r["status"], (1, 2, 6)).all() and (r["status"] == 6).sum() >= B // 2, r["status"]
    st = dict(zip(oracle_mod.STAT_COLUMNS, r["stats"].sum(axis=0)))
    assert st["resto_calls"] >= B // 2 and st["resto_iters"] > st["resto_calls"]
    assert np.isfinite(r["x"]).all() and np.all(r["x"] >= lbx) and np.all(r["x"] <= ubx)
    try:
        oracle_mod.set_option("resto", 0)
        r0 = oracle_mod.solve(sp, obs, p, x0, lbx, ubx, lbg, ubg)
    finally:
        oracle_mod.clear_options()
    assert np.isin(r0["status"], (1, 2)).all() and r0["stats"][:, 2].sum() == 0
    assert (r0["iters"] <= r["iters"]).all()


@pytest.mark.parametrize("source", ["constructed", "population"])
def test_infeasible_exit_is_a_local_minimiser_of_the_violation(pkg, oracle_mod, source):
    """Independent certificate for IPOPT's exit "Converged to a point of local infeasibility" (Infeasible_Problem_Detected):
    the returned w must be a first-order local minimiser of the l1 constraint violation subject to the control bounds.
    Checked with a different tool -- an LP (scipy / HiGHS) on the linearised rows in a box of radius 1e-2 around w,
        min sum(n + p)  s.t.  lbg <= g(w) + J dw + n - p <= ubg,  lbx <= w + dw <= ubx,  |dw| <= 1e-2,  n, p >= 0 --
    whose optimum may not undercut the violation at w itself (no direction decreases it), while the violation is far
    above the tolerance (so "infeasible" is the right verdict, not an artefact of the restoration phase).
    "constructed": NMPC_TT instances whose stage-0 rows are violated on purpose; "population": the exits that occur by
    themselves in a random Race Track 2 population (starts inside the keep-out disc of the script's first obstacle)."""
    from scipy.optimize import linprog
    if source == "constructed":
        sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, "nmpc_tt")
        B = 16
        p = _infeasible_batch(pkg, sc, B)
        x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
        need = B // 2
    else:
        sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, "race_track_2")
        B = 512
        p, _ = pkg.random_instances(sc, B, seed=5)
        x0 = np.zeros((B, sc.n_w))
        need = 8
    r = oracle_mod.solve(sp, obs, p, x0, lbx, ubx, lbg, ubg)
    n, m = sc.n_w, sc.n_g
    fu, fl = np.isfinite(ubg), np.isfinite(lbg)
    checked = 0
    for i in np.flatnonzero(r["status"] == 6)[:12]:
        ev = oracle_mod.evaluate(sp, obs, r["x"][i], p[i])          # g, J of the literal NLP (pinned to autograd in test_oracle_functions.py)
        g, J = ev["g"], ev["J"]
        theta = np.maximum(lbg - g, 0).sum() + np.maximum(g - ubg, 0).sum()
        assert theta > 1e-6, (i, theta)                             # (the constructed violations are 3e-6 ... 25)
        A = np.hstack([J, np.eye(m), -np.eye(m)])
        res = linprog(np.concatenate([np.zeros(n), np.ones(2 * m)]),
                      A_ub=np.vstack([A[fu], -A[fl]]), b_ub=np.concatenate([(ubg - g)[fu], (g - lbg)[fl]]),
                      bounds=[(max(lbx[k] - r["x"][i][k], -1e-2), min(ubx[k] - r["x"][i][k], 1e-2)) for k in range(n)] + [(0, None)] * (2 * m),
                      method="highs")
        assert res.status == 0
        assert theta - res.fun <= 1e-6 * theta + 1e-9, (i, theta, res.fun)
        checked += 1
    assert checked >= need


def test_restoration_elimination_matches_explicit_system(pkg, oracle_mod):
    """The restoration problem's Newton step with the n / p variables ELIMINATED (AugRestoSystemSolver: row weights
    Om = 1 / (1/D_s + 1/D_n + 1/D_p); what the CUDA kernel does) against the same algorithm factoring the condensed
    system with n, p as explicit variables (90 + 2 * 48 unknowns at N = 5): same statuses, same iteration counts, same
    returned points."""
    sc = pkg.SCENARIOS["nmpc_tt"].with_horizon(5)
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    lbx, ubx, lbg, ubg = sc.bounds()
    B = 24
    p = _infeasible_batch(pkg, sc, B, seed=5)
    p[::3] = pkg.random_instances(sc, B, seed=6)[0][::3]                     # a third of them ordinary instances
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    a = oracle_mod.solve(sp, sc.obstacle_table(), p, x0, lbx, ubx, lbg, ubg)
    try:
        oracle_mod.set_option("resto_explicit", 1)
        b = oracle_mod.solve(sp, sc.obstacle_table(), p, x0, lbx, ubx, lbg, ubg)
    finally:
        oracle_mod.clear_options()
    assert a["stats"][:, 2].sum() > 0                                       # restoration phases happened
    same = (a["status"] == b["status"]) & (a["iters"] == b["iters"])
    assert same.mean() >= 0.9, (a["status"], b["status"], a["iters"], b["iters"])
    assert np.abs(a["x"][same] - b["x"][same]).max() <= 1e-6
    assert np.abs(a["f"][same] - b["f"][same]).max() <= 1e-6 * np.abs(a["f"][same]).max()


def test_watchdog_and_soft_restoration_are_exercised(pkg, oracle_mod):
    """Closed-loop populations reach IPOPT's watchdog, soft restoration phase and filter resets; none of them may turn
    a solve that converges without them into a failure (they exist to help), and the counters say they ran."""
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, "t_trajectory")
    B = 192
    p, _ = pkg.random_instances(sc, B, seed=2000)
    x0 = np.zeros((B, sc.n_w))
    r = oracle_mod.solve(sp, obs, p, x0, lbx, ubx, lbg, ubg)
    st = dict(zip(oracle_mod.STAT_COLUMNS, r["stats"].sum(axis=0)))
    assert st["watchdog_starts"] > 0 and st["filter_resets"] > 0, st
    try:
        oracle_mod.set_option("watchdog_trigger", 0); oracle_mod.set_option("max_soft_resto", 0); oracle_mod.set_option("max_filter_resets", 0)
        r0 = oracle_mod.solve(sp, obs, p, x0, lbx, ubx, lbg, ubg)
    finally:
        oracle_mod.clear_options()
    assert (r["status"] == 0).sum() >= (r0["status"] == 0).sum() - 2
    both = (r["status"] == 0) & (r0["status"] == 0) & (r["stats"][:, 4:7].sum(axis=1) == 0)
    assert both.sum() > B // 2
    assert np.array_equal(r["iters"][both], r0["iters"][both])             # untouched instances follow the same iterates
    assert np.abs(r["f"][both] - r0["f"][both]).max() <= 1e-12 * np.abs(r["f"][both]).max()


@pytest.mark.parametrize("name", ["t_trajectory", "race_track_2", "nmpc_tt"])
def test_independent_solver_cross_check(pkg, oracle_mod, name):
    """Third-party check that IS available here (CasADi / IPOPT are not): scipy's SLSQP -- a different algorithm (SQP,
    active set) on the literal NLP (functions and first derivatives from torch.autograd of oracle/nlp_ref.py) --
    started near the oracle's converged golden solutions must return the same f* (rel 1e-7: the bound relaxation) and
    the same first input u0*: the oracle's points are local minimisers of the reference's NLP, not artefacts of its
    IPM.  u0* is compared to 1e-5 per component, except along directions in which the objective is flat to 1e-7, the accuracy SLSQP reaches (the
    gimbal yaw rate barely enters the cost over one stage -- SURVEY App. D.1 -- so different algorithms stop at
    different, equally optimal values there; the interior-point method returns the analytic centre of that face)."""
    from scipy.optimize import minimize
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod, name)
    G = np.load(GOLD / f"solves_{name}.npz")
    idx = [i for i in range(len(G["status"])) if G["status"][i] == 0][:6]
    assert idx
    fin_u, fin_l = np.isfinite(ubg), np.isfinite(lbg)
    rng = np.random.default_rng(1)
    checked = 0
    for i in idx:
        pi = G["p"][i]
        # f, g, grad f, J of the literal NLP from the oracle's evaluator (pinned to torch.autograd of oracle/nlp_ref.py by
        # test_oracle_functions.py); the SOLVER is the independent part of this test
        ev = lambda w: oracle_mod.evaluate(sp, obs, w, pi)

        def fg(w):
            d = ev(w)
            return d["f"], d["grad"]

        def cons(w):
            g = ev(w)["g"]
            return np.concatenate([(ubg - g)[fin_u], (g - lbg)[fin_l]])

        def cons_jac(w):
            J = ev(w)["J"]
            return np.concatenate([-J[fin_u], J[fin_l]])

        xs = G["x"][i]
        w0 = np.clip(xs + 1e-3 * (ubx - lbx) * rng.standard_normal(xs.size), lbx, ubx)
        res = minimize(fg, w0, jac=True, method="SLSQP", bounds=list(zip(lbx, ubx)),
                       constraints=[dict(type="ineq", fun=cons, jac=cons_jac)], options=dict(maxiter=400, ftol=1e-14))
        if res.status not in (0, 8):          # 8 = "positive directional derivative": SLSQP's usual exit at ftol = 1e-14
            continue
        if cons(res.x).min() < -1e-7:
            continue
        checked += 1
        # IPOPT's point sits on bounds relaxed by 1e-8 |b| and its f is evaluated after clipping: allow that much
        assert abs(res.fun - G["f"][i]) <= 1e-7 * abs(G["f"][i]) + 1e-8, (res.fun, G["f"][i])
        bad = np.abs(res.x[:6] - xs[:6]) > 1e-5 * np.abs(xs[:6]).max()
        if bad.any():
            w_sw = res.x.copy(); w_sw[:6][bad] = xs[:6][bad]                      # swap in the oracle's values of those components
            f_sw, _ = fg(w_sw)
            assert abs(f_sw - res.fun) <= 1e-7 * abs(res.fun) and cons(w_sw).min() >= -1e-6, (res.x[:6], xs[:6], f_sw - res.fun)
            assert bad.sum() <= 3
    assert checked >= 2


def test_warm_started_multipliers_same_solution_fewer_iterations(pkg, oracle_mod):
    """NON-REFERENCE mode: IPOPT's WarmStartIterateInitializer restated in the oracle_mod.  Re-solving an NLP from its own
    primal-dual solution takes far fewer iterations and returns the same objective; a NaN in lam_x0[b, 0] falls back to the
    cold start for that instance (bit-identical to the call without multiplier guesses)."""
    sc = pkg.SCENARIOS["t_trajectory"]
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    lbx, ubx, lbg, ubg = sc.bounds(); obs = sc.obstacle_table()
    B = 24
    p, _ = pkg.random_instances(sc, B, seed=11)
    r0 = oracle_mod.solve(sp, obs, p, np.zeros((B, sc.n_w)), lbx, ubx, lbg, ubg)
    cold = oracle_mod.solve(sp, obs, p, r0["x"], lbx, ubx, lbg, ubg)
    lx = r0["lam_x"].copy(); lx[::4, 0] = np.nan
    warm = oracle_mod.solve(sp, obs, p, r0["x"], lbx, ubx, lbg, ubg, lam_x0=lx, lam_g0=r0["lam_g"])
    ok = (r0["status"] == 0) & (cold["status"] == 0) & (warm["status"] == 0)
    assert ok.mean() > 0.9
    assert np.array_equal(warm["x"][::4], cold["x"][::4]) and np.array_equal(warm["iters"][::4], cold["iters"][::4])
    w = ok.copy(); w[::4] = False
    assert warm["iters"][w].mean() < 0.6 * cold["iters"][w].mean()
    # the NLP is non-convex: a re-solve may settle in another local solution; nearly all return the same objective
    assert np.mean(np.abs(warm["f"][w] - cold["f"][w]) <= 1e-6 * np.maximum(1.0, np.abs(cold["f"][w]))) >= 0.9


# ---- second, structurally independent restatement of IPOPT's main loop (oracle/ipm_fullspace.py) -----------------------
# (scenario, source, index): source "golden" = tests/golden/solves_<scenario>.npz, ("random", seed, B) = cold-started
# instance `index` of scenarios.random_instances(scenario, B, seed)
FULLSPACE_CASES = [("t_trajectory", "golden", 0),      # the T = 0.2 scripts' own cold first solve (f* = 248.10109)
                   ("nmpc_tt", "golden", 0),           # NMPC_TT.py's own cold first solve: 100 iterations, Maximum_Iterations_Exceeded
                   ("nmpc_tt", "golden", 4),           # warm-started T = 1 solve
                   ("nmpc_tt", "golden", 10),          # 68 iterations, one accepted second-order correction
                   ("race_track_2", "golden", 1),      # ten obstacle rows: 240 rows, 570 x 570 full-space system
                   ("10_obstacles", "golden", 2),
                   ("gimbal_less", "golden", 0),       # the other NLP (MATLAB/Dynamic Obstacles/NMPC_TT.m): 45 variables, 32 rows, its script's first solve
                   ("t_trajectory", ("random", 2000, 192), 81),    # six watchdog phases in 100 iterations
                   ("t_trajectory", ("random", 3, 1024), 388)]     # four watchdog phases, a filter reset, two soft restoration
                                                                   # steps, then the restoration phase is entered (iteration 98)


def _fullspace_case(args):
    """Worker (own process): one instance through the C++ oracle (with its iteration log) and through the full-space
    restatement."""
    name, source, idx = args
    import sys
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import torch
    torch.set_num_threads(1)
    import b200nmpc
    import oracle
    from oracle import ipm_fullspace
    sc = b200nmpc.SCENARIOS[name]
    five = getattr(sc, "model", 0) == 1
    sp = oracle.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov, model=1 if five else 0)
    obs = np.zeros((0, 3)) if five else sc.obstacle_table()
    lbx, ubx, lbg, ubg = sc.bounds()
    if source == "golden":
        G = np.load(GOLD / f"solves_{name}.npz")
        p, x0 = G["p"][idx], G["x0"][idx]
    else:
        p = b200nmpc.random_instances(sc, source[2], seed=source[1])[0][idx]
        x0 = np.zeros(sc.n_w)
    lg = oracle.solve_log(sp, obs, p, x0, lbx, ubx, lbg, ubg)
    full = oracle.solve(sp, obs, p[None], x0[None], lbx, ubx, lbg, ubg)
    rs = nlp_ref.RefSpec5(sc.T, sc.N) if five else nlp_ref.RefSpec(T=sc.T, N=sc.N, obstacles=sc.obstacles, uav_r=sc.uav_r, w1=sc.w1, w2=sc.w2)
    log, events = [], []
    q = ipm_fullspace.solve(ipm_fullspace.Problem(rs, p), x0, lbx, ubx, lbg, ubg, log=log, events=events)
    return dict(name=name, idx=idx, o_status=int(lg["status"]), o_iters=int(lg["iters"]), o_log=lg["log"], o_x=lg["x"], o_f=lg["f"],
                o_lam_x=full["lam_x"][0], o_lam_g=full["lam_g"][0], o_stats=full["stats"][0],
                q_status=q["status"], q_iters=q["iters"], q_log=np.array(log), q_x=q.get("x"), q_f=q.get("f"),
                q_lam_x=q.get("lam_x"), q_lam_g=q.get("lam_g"), q_counters=q["counters"], q_events=events)


def test_fullspace_ipm_reproduces_oracle_iterates(pkg, oracle_mod, monkeypatch):
    """The C++ oracle (condensed n_w x n_w Cholesky, AD jets) against oracle/ipm_fullspace.py (IPOPT's full-space
    augmented system with a Bunch-Kaufman inertia count, torch.autograd derivatives of the literal formulas, numpy
    state): same return status, same iteration count, and the same iteration LOG -- barrier parameter, inertia
    perturbation delta_w, primal and dual step lengths, number of trial points per line search, primal / dual
    infeasibility -- on every iteration, including the 100 wandering iterations of NMPC_TT.py's own cold first solve; same
    solution and multipliers.  Two instances walk through IPOPT's globalisation heuristics: watchdog phases start at
    the same iterations and end the same way, the filter is reset at the same iteration, the soft restoration steps
    have the same lengths, and the restatement asks for the restoration phase (which it does not contain) exactly
    where the oracle enters it.  What this pins: the elimination of slacks and multipliers, 'reduced matrix positive
    definite <=> inertia (n + m, m, 0)', the hand / jet derivatives, the sign and scaling conventions of lam_x / lam_g,
    and the bookkeeping of filter line search, watchdog and soft restoration -- by an implementation that shares none of
    the code.  What it cannot pin: a misreading of IPOPT common to both restatements (parity stays unpinned, DESIGN.md
    section 5)."""
    import concurrent.futures as cf
    import multiprocessing as mp
    from oracle import STATUS_NAMES
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):      # one BLAS thread per worker: same summation order every run
        monkeypatch.setenv(v, "1")
    with cf.ProcessPoolExecutor(max_workers=min(len(FULLSPACE_CASES), mp.cpu_count()), mp_context=mp.get_context("spawn")) as ex:
        results = list(ex.map(_fullspace_case, FULLSPACE_CASES))
    for r in results:
        tag = (r["name"], r["idx"])
        main_rows = r["o_log"][:, 8] < 1000                                  # (rows >= 1000 are iterations of the restoration phase)
        n_main = int(np.argmin(main_rows)) if not main_rows.all() else len(main_rows)
        L, M = r["o_log"][:n_main, :8], r["q_log"]
        st = r["o_stats"]
        if st[2] > 0:          # the oracle entered the restoration phase: the restatement must stop right there
            assert r["q_status"] == "needs_restoration" and r["q_iters"] == n_main == len(M), (tag, r["q_status"], r["q_iters"], n_main)
        else:
            assert r["q_status"] == STATUS_NAMES[r["o_status"]], (tag, r["q_status"], r["o_status"])
            assert r["q_iters"] == r["o_iters"] == len(M), (tag, r["q_iters"], r["o_iters"])
        assert (r["q_counters"]["watchdog_starts"], r["q_counters"]["filter_resets"]) == (st[4], st[6]), (tag, r["q_counters"], st)
        assert r["q_counters"]["soft_resto_steps"] == st[5] or st[2] > 0, (tag, r["q_counters"], st)
        assert L.shape == M.shape, tag
        wandering = r["o_iters"] >= 60                                     # rounding differences grow along a long non-converging run
        rt = (3e-2 if st[5] > 0 else 1e-3) if wandering else 1e-6            # (measured: 2e-3 resp. 1e-6 ... 1e-5)
        assert np.array_equal(L[:, 7], M[:, 7]), (tag, "trial points per line search")
        assert np.allclose(L[:, 0], M[:, 0], rtol=1e-12, atol=0), (tag, "mu")
        assert np.allclose(L[:, 4], M[:, 4], rtol=1e-9, atol=0), (tag, "delta_w")
        assert np.allclose(L[:, 5:7], M[:, 5:7], rtol=rt, atol=0), (tag, "alpha_pr / alpha_du")
        assert np.allclose(L[:, 1], M[:, 1], rtol=1e-7 if wandering else 1e-9, atol=0), (tag, "objective")
        assert np.allclose(L[:, 2], M[:, 2], rtol=rt, atol=1e-11), (tag, "inf_pr")
        assert np.allclose(L[:, 3], M[:, 3], rtol=10 * rt, atol=1e-11), (tag, "inf_du")
        if st[2] == 0:
            assert abs(r["q_f"] - r["o_f"]) <= (1e-8 if wandering else 1e-11) * abs(r["o_f"]), tag
            assert np.abs(r["q_x"] - r["o_x"]).max() <= (1e-5 if r["o_status"] else 1e-8), tag
        if r["o_status"] == 0:
            assert np.abs(r["q_lam_x"] - r["o_lam_x"]).max() <= 1e-7 * max(1.0, np.abs(r["o_lam_x"]).max()), tag
            assert np.abs(r["q_lam_g"] - r["o_lam_g"]).max() <= 1e-7 * max(1.0, np.abs(r["o_lam_g"]).max()), tag
    tags = lambda r: "".join(chr(int(t) % 1000) for t in r["o_log"][:, 8])
    assert any("H" in tags(r) for r in results)                                                    # an accepted second-order correction
    assert any(r["o_status"] == 1 for r in results) and any(r["o_log"][:, 4].max() > 0 for r in results)   # max_iter run; inertia corrections
    assert sum(r["q_counters"]["watchdog_starts"] for r in results) >= 10                          # watchdog phases ...
    assert any(r["q_counters"]["filter_resets"] for r in results) and any(r["q_counters"]["soft_resto_steps"] for r in results)


def test_fullspace_fixture_matches_oracle(pkg, oracle_mod):
    """tests/golden/fullspace_solves.npz holds solutions computed by the independent full-space restatement
    (tests/golden/make_fullspace_golden.py); the C++ oracle returns the same status, iteration count, f*, x* and multipliers on
    those instances (the GPU suite compares the kernel with the same file)."""
    F = np.load(GOLD / "fullspace_solves.npz")
    for k, (name, idx) in enumerate(zip(F["names"], F["idx"])):
        name = str(name)
        sc = pkg.SCENARIOS[name]
        sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov, model=sc.model)
        obs = np.zeros((0, 3)) if sc.model == 1 else sc.obstacle_table()
        G = np.load(GOLD / f"solves_{name}.npz")
        r = oracle_mod.solve(sp, obs, G["p"][idx][None], G["x0"][idx][None], *sc.bounds())
        assert r["status"][0] == F["status"][k] and r["iters"][0] == F["iters"][k], (name, idx)
        conv = F["status"][k] == 0
        assert abs(r["f"][0] - F["f"][k]) <= (1e-11 if conv else 1e-7) * abs(F["f"][k]), (name, idx)
        assert np.abs(r["x"][0] - F[f"x_{k}"]).max() <= (1e-8 if conv else 1e-4), (name, idx)
        if conv:
            assert np.abs(r["lam_g"][0] - F[f"lam_g_{k}"]).max() <= 1e-7 * max(1.0, np.abs(F[f"lam_g_{k}"]).max()), (name, idx)
            assert np.abs(r["lam_x"][0] - F[f"lam_x_{k}"]).max() <= 1e-7 * max(1.0, np.abs(F[f"lam_x_{k}"]).max()), (name, idx)
