"""GPU (-m gpu): the CUDA path, called through the C ABI (ctypes), against the CPU oracle and the golden
fixtures.  Tolerances are the north star's: first input u0* rel 1e-6, objective rel 1e-8, identical active set."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).resolve().parent / "golden"
NAMES = ["nmpc_tt", "t_trajectory", "plus_trajectory", "race_trajectory_1", "race_track_2", "10_obstacles"]
U0_RTOL, F_RTOL = 1e-6, 1e-8


def _active_mismatch(vr, vg, lo, hi, strict=1e-7, loose=1e-5):
    """Active-set comparison with hysteresis: a bound counts as differently active only when one solution sits
    on it (distance < strict) while the other is clearly off it (distance > loose).  Interior-point solutions
    keep active slacks at ~mu/lambda ~ 1e-9, so strongly active bounds are far inside `strict`."""
    bad = np.zeros(vr.shape, dtype=bool)
    for br, bg in ((vr - lo, vg - lo), (hi - vr, hi - vg)):
        with np.errstate(invalid="ignore"):
            bad |= ((br < strict) & (bg > loose)) | ((bg < strict) & (br > loose))
    return bad


def _compare(ref, got, status_g, iters_g, bounds, need_all_status=True):
    lbx, ubx, lbg, ubg = bounds
    if need_all_status:
        assert np.array_equal(ref["status"], status_g)
    both = (ref["status"] == 0) & (status_g == 0)
    assert both.sum() == (ref["status"] == 0).sum()
    rf = np.abs(ref["f"] - got["f"]) / np.maximum(1.0, np.abs(ref["f"]))
    assert rf[both].max() <= F_RTOL
    u0r, u0g = ref["x"][:, :6], got["x"][:, :6]
    ru = np.abs(u0r - u0g).max(axis=1) / np.maximum(1e-12, np.abs(u0r).max(axis=1))
    assert ru[both].max() <= U0_RTOL
    assert not _active_mismatch(ref["x"][both], got["x"][both], lbx, ubx).any()       # identical active set: controls
    assert not _active_mismatch(ref["g"][both], got["g"][both], lbg, ubg).any()       # identical active set: g rows
    return both


def _compare_bulk(ref, got, bounds, frac=0.995):
    """Population version of _compare for closed-loop populations (thousands of solves, a few of them ill-conditioned):
    at least `frac` of the instances where both sides converged meet the north star's tolerances (u0* 1e-6, f* 1e-8,
    identical active set); every one of them agrees in f* to 1e-6 (two implementations never return different local
    solutions).  The rest are instances whose first input has a direction along which the objective is flat to
    rounding (weakly determined gimbal rates): f* agrees, u0* only to ~1e-4."""
    lbx, ubx, lbg, ubg = bounds
    n = len(ref["f"])
    if n == 0:
        return 0.0
    rf = np.abs(ref["f"] - got["f"]) / np.maximum(1.0, np.abs(ref["f"]))
    u0r, u0g = ref["x"][:, :6], got["x"][:, :6]
    ru = np.abs(u0r - u0g).max(axis=1) / np.maximum(1e-12, np.abs(u0r).max(axis=1))
    act = _active_mismatch(ref["x"], got["x"], lbx, ubx).any(axis=1) | _active_mismatch(ref["g"], got["g"], lbg, ubg).any(axis=1)
    good = (rf <= F_RTOL) & (ru <= U0_RTOL) & ~act
    assert good.mean() >= frac, (good.mean(), np.sort(ru)[-5:], np.sort(rf)[-5:], int(act.sum()))
    assert rf.max() <= 1e-6, np.sort(rf)[-5:]
    return good.mean()


@pytest.mark.parametrize("name", NAMES)
def test_casadi_records(pkg, name):
    """PINNING HOOK: real CasADi + IPOPT records (tests/golden/casadi_<name>.npz, made by bench/run_casadi.py where CasADi
    is installed) take precedence over the oracle's own fixtures: same status, f* 1e-8, u0* 1e-6 against REAL IPOPT."""
    f = GOLD / f"casadi_{name}.npz"
    if not f.exists():
        pytest.skip("no real-reference records in this repository (CasADi unavailable in the build image): parity unpinned")
    C = np.load(f, allow_pickle=False)
    sc = pkg.SCENARIOS[name]
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=len(C["f"]))
    lbx, ubx, lbg, ubg = sc.bounds()
    sol = s(x0=C["x0"], p=C["p"], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    stg = s.stats()["return_status"]
    assert np.array_equal(C["status"], stg), (C["status"], stg)
    ok = C["status"] == 0
    # (a uniform f* mismatch of 1 ... 2e-8 would point at the evaluation point of sol['f'], DESIGN.md section 5 "known candidate deviation")
    assert np.all(np.abs(sol["f"][ok] - C["f"][ok]) <= F_RTOL * np.abs(C["f"][ok]))
    assert np.all(np.abs(sol["x"][ok, :6] - C["x"][ok, :6]).max(axis=1) <= U0_RTOL * np.abs(C["x"][ok, :6]).max(axis=1))


def test_fullspace_fixtures(pkg):
    """The kernel against solutions computed by the INDEPENDENT full-space restatement of IPOPT's algorithm
    (oracle/ipm_fullspace.py -- full-space KKT system, LDL^T inertia count, autograd derivatives; fixture
    tests/golden/fullspace_solves.npz made by tests/golden/make_fullspace_golden.py), i.e. against an implementation that
    shares no code with the C++ oracle the other tests use: same return status, f* rel 1e-8, u0* rel 1e-6, multipliers."""
    F = np.load(GOLD / "fullspace_solves.npz")
    same_iters = 0
    for k, (name, idx) in enumerate(zip(F["names"], F["idx"])):
        name = str(name)
        sc = pkg.SCENARIOS[name]
        G = np.load(GOLD / f"solves_{name}.npz")
        s = pkg.nlpsol("solver", "ipm", sc, max_batch=1)
        lbx, ubx, lbg, ubg = sc.bounds()
        sol = s(x0=G["x0"][idx][None], p=G["p"][idx][None], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
        st = s.stats()
        assert int(st["return_status"][0]) == int(F["status"][k]), (name, idx, st["return_status"], F["status"][k])
        same_iters += int(st["iter_count"][0]) == int(F["iters"][k])
        if F["status"][k] == 0:
            x = F[f"x_{k}"]
            assert abs(sol["f"][0] - F["f"][k]) <= F_RTOL * abs(F["f"][k]), (name, idx)
            nu = sc.nu
            assert np.abs(sol["x"][0][:nu] - x[:nu]).max() <= U0_RTOL * np.abs(x[:nu]).max(), (name, idx)
            lg = F[f"lam_g_{k}"]
            assert np.abs(sol["lam_g"][0] - lg).max() <= 1e-3 * max(1.0, np.abs(lg).max()), (name, idx)
    assert same_iters >= len(F["idx"]) - 2, same_iters


@pytest.mark.parametrize("name", NAMES)
def test_golden_fixtures(pkg, name):
    """Committed oracle solutions (first closed-loop steps of each reference script + seeded instances)."""
    sc = pkg.SCENARIOS[name]
    G = np.load(GOLD / f"solves_{name}.npz")
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=len(G["f"]))
    lbx, ubx, lbg, ubg = sc.bounds()
    sol = s(x0=G["x0"], p=G["p"], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    st = s.stats()
    ref = {k: G[k] for k in ("x", "f", "g", "status", "iters")}
    stg = st["return_status"]
    assert np.array_equal(G["status"], stg), (G["status"], stg)            # no status flips on any fixture
    both = (G["status"] == 0) & (stg == 0)
    sel = lambda d: {k: d[k][both] for k in ("x", "f", "g")}
    r2 = sel(ref); r2["status"] = G["status"][both]
    _compare(r2, sel(sol), stg[both], st["iter_count"][both], (lbx, ubx, lbg, ubg))
    okr = G["status"] == 0                               # solution parity wherever the oracle converged
    assert (np.abs(sol["f"][okr] - G["f"][okr]) <= 1e-7 * np.abs(G["f"][okr])).all()
    assert (np.abs(sol["x"][okr, :6] - G["x"][okr, :6]).max(axis=1) <= 1e-5 * np.abs(G["x"][okr, :6]).max(axis=1)).all()
    # the two implementations follow the same iterates: iteration counts agree on (nearly) every instance
    assert (st["iter_count"][both] == G["iters"][both]).mean() >= 0.85
    assert np.abs(sol["lam_g"][both] - G["lam_g"][both]).max() <= 1e-6 * max(1.0, np.abs(G["lam_g"][both]).max())


@pytest.mark.parametrize("name,N,B", [("t_trajectory", 15, 96), ("nmpc_tt", 15, 64), ("race_track_2", 15, 64),
                                       ("10_obstacles", 15, 48), ("race_track_2", 30, 32), ("t_trajectory", 5, 16)])
def test_solve_parity_random(pkg, oracle_mod, name, N, B):
    sc = pkg.SCENARIOS[name]
    if N != sc.N:
        sc = sc.with_horizon(N)
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=2024 + N)
    rng = np.random.default_rng(5)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N)) + 0.01 * rng.standard_normal((B, sc.n_w))
    x0[: B // 4] = 0.0                      # cold starts like the scripts' first step (NMPC_TT.py:329)
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    ref = oracle_mod.solve(sp, sc.obstacle_table(), p, x0, lbx, ubx, lbg, ubg)
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    sol = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    st = s.stats()
    # identical algorithm, rounding-level differences only
    assert (ref["status"] == st["return_status"]).mean() >= 0.95
    both = (ref["status"] == 0) & (st["return_status"] == 0)
    assert both.sum() >= 0.95 * (ref["status"] == 0).sum()
    okr = ref["status"] == 0                               # solution parity wherever the oracle converged
    assert (np.abs(sol["f"][okr] - ref["f"][okr]) <= 1e-7 * np.abs(ref["f"][okr])).all()
    sel = lambda d: {k: (v[both] if v is not None else None) for k, v in d.items() if k in ("x", "f", "g", "status")}
    r2 = sel(ref); r2["status"] = ref["status"][both]
    _compare(r2, sel(sol), st["return_status"][both], st["iter_count"][both], (lbx, ubx, lbg, ubg))


def test_per_instance_obstacles(pkg, oracle_mod):
    """Monte-Carlo obstacle fields (BASELINE config 4): obst [B][n_obs][3]."""
    sc = pkg.SCENARIOS["10_obstacles"]
    B = 24
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=9)
    rng = np.random.default_rng(10)
    obs = np.tile(sc.obstacle_table(), (B, 1, 1))
    obs[:, :3, :2] += rng.uniform(-100, 100, (B, 3, 2))
    obs[:, 0, :2] = p[:, :2] + np.array([260.0, 40.0])          # one obstacle ahead of every UAV
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs)
    ref = oracle_mod.solve(sp, obs, p, x0, lbx, ubx, lbg, ubg, obs_per_instance=True)
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    sol = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, obstacles=obs)
    st = s.stats()
    # two instances of this batch are hard (94 / 100 iterations in the oracle): their final status may flip
    assert (ref["status"] == st["return_status"]).mean() >= 0.9
    both = (ref["status"] == 0) & (st["return_status"] == 0)
    assert both.sum() >= 20
    sel = lambda d: {k: d[k][both] for k in ("x", "f", "g")}
    r2 = sel(ref); r2["status"] = ref["status"][both]
    _compare(r2, sel(sol), st["return_status"][both], st["iter_count"][both], (lbx, ubx, lbg, ubg))


def test_per_instance_weights(pkg, oracle_mod):
    """cost weights (w1, w2) as per-instance inputs (the reference's RL-style weight sweeps)."""
    sc = pkg.SCENARIOS["t_trajectory"]
    B = 36
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=31)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    pairs = [(1.0, 2.0), (0.5, 4.0), (3.0, 0.5)]
    wts = np.array([pairs[i % 3] for i in range(B)])
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    sol = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, weights=wts)
    st = s.stats()
    for k, (w1, w2) in enumerate(pairs):
        sel = np.arange(B) % 3 == k
        sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, w1, w2, sc.vfov, sc.hfov)
        ref = oracle_mod.solve(sp, sc.obstacle_table(), p[sel], x0[sel], lbx, ubx, lbg, ubg)
        got = {q: sol[q][sel] for q in ("x", "f", "g")}
        assert (ref["status"] == st["return_status"][sel]).mean() >= 0.9
        both = (ref["status"] == 0) & (st["return_status"][sel] == 0)
        assert both.sum() >= 8
        r2 = {q: ref[q][both] for q in ("x", "f", "g")}; r2["status"] = ref["status"][both]
        _compare(r2, {q: got[q][both] for q in got}, st["return_status"][sel][both], st["iter_count"][sel][both], (lbx, ubx, lbg, ubg))
        # function level with the same weights
        ev = s.evaluate(sol["x"][sel], p[sel], weights=wts[sel])
        for j in range(3):
            fo = oracle_mod.evaluate(sp, sc.obstacle_table(), sol["x"][sel][j], p[sel][j])
            assert abs(float(ev["f"][j]) - fo["f"]) <= 1e-12 * abs(fo["f"])
            assert np.allclose(ev["grad"][j].cpu().numpy(), fo["grad"], rtol=1e-8, atol=1e-9 * np.abs(fo["grad"]).max())
    # the weights applied to that call only: the next call uses the spec's (1, 2) again
    sol2 = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    k0 = np.arange(B) % 3 == 0
    ok = (st["return_status"] == 0) & (s.stats()["return_status"] == 0) & k0
    assert np.allclose(sol2["f"][ok], sol["f"][ok], rtol=1e-9)
    assert not np.allclose(sol2["f"][~k0], sol["f"][~k0], rtol=1e-3)


def test_predicted_target_trajectory(pkg, oracle_mod):
    """SURVEY 8f-2 / north star "p = [UAV state; predicted target trajectory]": per-stage target positions."""
    sc = pkg.SCENARIOS["t_trajectory"]
    B = 24
    lbx, ubx, lbg, ubg = sc.bounds()
    p, vw = pkg.random_instances(sc, B, seed=41)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    k = np.arange(sc.N)[None, :, None]
    head = np.stack([np.cos(p[:, 10] + 0.0), np.sin(p[:, 10])], axis=1)[:, None, :]
    traj = p[:, None, 8:10] + sc.T * k * vw[:, None, 0:1] * head            # constant-velocity prediction [B, N, 2]
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    ref = oracle_mod.solve(sp, sc.obstacle_table(), p, x0, lbx, ubx, lbg, ubg, target_traj=traj)
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    sol = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, target_traj=traj)
    st = s.stats()
    assert (ref["status"] == st["return_status"]).mean() >= 0.9
    both = (ref["status"] == 0) & (st["return_status"] == 0)
    assert both.sum() >= 18
    r2 = {q: ref[q][both] for q in ("x", "f", "g")}; r2["status"] = ref["status"][both]
    _compare(r2, {q: sol[q][both] for q in ("x", "f", "g")}, st["return_status"][both], st["iter_count"][both], (lbx, ubx, lbg, ubg))
    ev = s.evaluate(sol["x"], p, target_traj=traj)
    for j in range(3):
        fo = oracle_mod.evaluate(sp, sc.obstacle_table(), sol["x"][j], p[j], target_traj=traj[j])
        assert abs(float(ev["f"][j]) - fo["f"]) <= 1e-12 * abs(fo["f"])
        assert np.allclose(ev["grad"][j].cpu().numpy(), fo["grad"], rtol=1e-8, atol=1e-9 * np.abs(fo["grad"]).max())
    # the prediction changes the problem (a moving target is not the frozen one) and applies to that call only
    base = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    assert not np.allclose(base["f"], sol["f"], rtol=1e-3)
    frozen = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, target_traj=np.tile(p[:, None, 8:10], (1, sc.N, 1)))
    assert np.array_equal(frozen["x"], base["x"]) and np.array_equal(frozen["f"], base["f"])
    # closed loop with the target's own prediction: tracks at least as well as the frozen-target loop of the reference
    from mpc_implementation_b200.closed_loop import ClosedLoop
    errs = []
    for pred in (False, True):
        cl = ClosedLoop(pkg.nlpsol("solver", "ipm", sc, max_batch=B), sc, p, target_vw=vw, predict_target=pred)
        if pred:      # stage 0 of the prediction is the current target
            assert torch.equal(cl.target_prediction()[:, 0, :], cl.p[:, 8:10])
        for _ in range(25):
            cl.step()
        errs.append(float(cl.err_sum.mean()))
    assert errs[1] <= 1.05 * errs[0], errs


def test_moving_obstacles_closed_loop(pkg):
    """Obstacle fields as data that move between steps (MATLAB/Dynamic Obstacles/Dynamic Obstacle avoidance.m:128-133):
    a slowly moving disc is placed so that it overlaps the path the unobstructed loop flies; the loop keeps its distance."""
    from mpc_implementation_b200.closed_loop import ClosedLoop
    sc = pkg.SCENARIOS["t_trajectory"]
    B, K, K0 = 6, 90, 45
    p0 = np.tile(np.array(list(sc.x_init) + list(sc.target_init)), (B, 1))
    p0[:, 1] += np.linspace(-6, 6, B); p0[:, 9] = p0[:, 1]                      # parallel lanes
    vw = np.tile(np.array([12.0, 0.0]), (B, 1))                                   # target straight along +x
    v_obs = np.array([-2.0, 0.0])

    def fly(obs, vel):
        s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
        cl = ClosedLoop(s, sc, p0, target_vw=vw, obstacles=obs, obstacle_vel=vel)
        path, conv = [], 0
        for k in range(K):
            cl.step(); conv += int(s.stats()["success"].sum())
            path.append(cl.p[:, :2].cpu().numpy().copy())
        return np.array(path), conv, cl

    free, _, _ = fly(None, None)                                                  # [K, B, 2]
    # disc of radius 35 whose centre passes 27 m beside the free path at step K0 (the track cuts 8 m into it)
    c_at = lambda k: free[K0] + np.array([0.0, 27.0]) + v_obs * sc.T * (k - K0)
    obs = np.tile(sc.obstacle_table(), (B, 1, 1)); obs[:, 0, :2] = c_at(-1); obs[:, 0, 2] = 35.0
    vel = np.zeros((B, sc.n_obs, 2)); vel[:, 0, :] = v_obs
    path, conv, cl = fly(obs, vel)
    assert torch.allclose(cl.obstacles[:, 0, :2], torch.as_tensor(c_at(K - 1), device=cl.obstacles.device), atol=1e-9)
    assert conv >= 0.6 * K * B, conv
    d_with = np.array([np.linalg.norm(path[k] - c_at(k), axis=1) for k in range(K)]).min(axis=0)
    d_free = np.array([np.linalg.norm(free[k] - c_at(k), axis=1) for k in range(K)]).min(axis=0)
    assert (d_free < 30.0).all(), d_free                          # it was in the way ...
    assert (d_with >= 35.0 - 1e-2).all(), d_with                  # ... and is avoided (r_uav + r_obs = 35)


def test_schedule_phases(pkg):
    """BASELINE config 3: instances of one batch start at different phases of the script's target schedule
    (T_Trajectory.py's con_t keyed on mpc_iter); phase 0 is the script's own loop."""
    from mpc_implementation_b200.closed_loop import ClosedLoop
    sc = pkg.SCENARIOS["t_trajectory"]
    p0 = np.array(list(sc.x_init) + list(sc.target_init))
    one = ClosedLoop(pkg.nlpsol("a", "ipm", sc), sc, p0)
    many = ClosedLoop(pkg.nlpsol("b", "ipm", sc, max_batch=3), sc, np.tile(p0, (3, 1)), phase=[0, 40, 400])
    for _ in range(30):
        one.step(); many.step()
    torch.cuda.synchronize()
    assert torch.equal(one.p[0], many.p[0]) and torch.equal(one.err_sum[0], many.err_sum[0])
    # the schedule is looked up on the device (nmpc_set_schedule): the target of the instance with phase 400 followed
    # schedule(400 + i), i = 0..29 -- replay the target's Euler steps (NMPC_TT.py:24-29) on the host
    for b, ph in ((1, 40), (2, 400)):
        tg = np.array(sc.target_init, dtype=np.float64)
        for i in range(30):
            v, w = sc.schedule(i + ph)
            tg = tg + sc.T * np.array([v * np.cos(tg[2]), v * np.sin(tg[2]), w])
        assert np.allclose(many.p[b, 8:].cpu().numpy(), tg, rtol=0, atol=1e-10), (b, many.p[b, 8:], tg)
    assert sc.schedule(429) != sc.schedule(29) and not torch.equal(many.p[0, 8:], many.p[2, 8:])
    # the unfused path (solve, then nmpc_step with the gathered (v, omega)) follows the same schedule
    two = ClosedLoop(pkg.nlpsol("c", "ipm", sc, max_batch=3), sc, np.tile(p0, (3, 1)), phase=[0, 40, 400])
    for _ in range(30):
        two.step(want_g=True)
    torch.cuda.synchronize()
    assert torch.equal(two.p, many.p) and torch.equal(two.err_sum, many.err_sum)
    v2, w2 = sc.schedule(429)
    assert float(two.vw[2, 0]) == v2 and float(two.vw[2, 1]) == w2


def test_function_level(pkg, oracle_mod):
    """nmpc_eval (f, g, grad f, J^T lam, Hess_L v) vs the oracle's dense derivatives."""
    for name in ("nmpc_tt", "race_track_2"):
        sc = pkg.SCENARIOS[name]
        B = 32
        rng = np.random.default_rng(3)
        lbx, ubx, _, _ = sc.bounds()
        p, _ = pkg.random_instances(sc, B, 4)
        w = lbx + rng.random((B, sc.n_w)) * (ubx - lbx)
        lam = rng.standard_normal((B, sc.n_g)); v = rng.standard_normal((B, sc.n_w))
        s = pkg.nlpsol("solver", "ipm", sc)
        r = {k: t.cpu().numpy() for k, t in s.evaluate(w, p, lam=lam, v=v, sigma=0.7).items()}
        sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs)
        for b in range(B):
            o = oracle_mod.evaluate(sp, sc.obstacle_table(), w[b], p[b], lam[b], 0.7, hessian=True)
            rel = lambda a, c: np.abs(a - c).max() / max(1.0, np.abs(c).max())
            assert rel(r["f"][b], o["f"]) <= 1e-13 and rel(r["g"][b], o["g"]) <= 1e-13
            assert rel(r["grad"][b], o["grad"]) <= 1e-12 and rel(r["jtv"][b], o["J"].T @ lam[b]) <= 1e-12
            assert rel(r["hv"][b], o["H"] @ v[b]) <= 1e-11


def test_host_and_device_entry_points_agree(pkg):
    """nmpc_solve (device pointers, torch stream) and nmpc_solve_host (numpy) give bit-identical results; so do
    different batch compositions (instances are independent: sharded == unsharded)."""
    sc = pkg.SCENARIOS["t_trajectory"]
    B = 40
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=31)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    host = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    dev = s(x0=torch.from_numpy(x0).cuda(), p=torch.from_numpy(p).cuda(), lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    torch.cuda.synchronize()
    assert np.array_equal(host["x"], dev["x"].cpu().numpy()) and np.array_equal(host["f"], dev["f"].cpu().numpy())
    perm = np.random.default_rng(0).permutation(B)[:17]
    sub = s(x0=x0[perm], p=p[perm], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    assert np.array_equal(sub["x"], host["x"][perm]) and np.array_equal(sub["lam_g"], host["lam_g"][perm])
    one = s(x0=x0[3], p=p[3], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)          # CasADi-style single call
    assert one["x"].shape == (sc.n_w,) and np.array_equal(one["x"], host["x"][3])


def test_closed_loop_matches_oracle_teacher_forced(pkg, oracle_mod):
    """First 25 steps of T_Trajectory.py's loop on the GPU (solve + nmpc_step); at every step the oracle solves
    the GPU's own (p, x0) (teacher forcing) and the shift is checked against the restated
    shift_timestep (NMPC_TT.py:13-30)."""
    from mpc_implementation_b200.closed_loop import ClosedLoop
    from oracle import nlp_ref
    sc = pkg.SCENARIOS["t_trajectory"]
    s = pkg.nlpsol("solver", "ipm", sc)
    cl = ClosedLoop(s, sc, np.array(list(sc.x_init) + list(sc.target_init)))
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs)
    lbx, ubx, lbg, ubg = sc.bounds()
    err = 0.0
    for i in range(25):
        p = cl.p.cpu().numpy()[0].copy(); w0 = cl.u_warm.cpu().numpy()[0].copy()
        sol = cl.step()
        ref = oracle_mod.solve(sp, sc.obstacle_table(), p, w0, lbx, ubx, lbg, ubg, nthreads=1)
        x = sol["x"].cpu().numpy()[0]
        assert ref["status"][0] == int(s.stats()["return_status"][0].item())
        assert abs(ref["f"][0] - float(sol["f"][0])) <= F_RTOL * abs(ref["f"][0])
        assert np.abs(ref["x"][0][:6] - x[:6]).max() <= U0_RTOL * np.abs(ref["x"][0][:6]).max()
        x0n, u0n, xsn = nlp_ref.shift_timestep(sc.T, p[:8], x.reshape(sc.N, 6).T, p[8:], sc.schedule(i))
        pn = cl.p.cpu().numpy()[0]
        assert np.allclose(pn[:8], x0n, rtol=0, atol=1e-12) and np.allclose(pn[8:], xsn, rtol=0, atol=1e-12)
        assert np.array_equal(cl.u_warm.cpu().numpy()[0], u0n.T.reshape(-1))
        xe, ye = nlp_ref.fov_centre(x0n)
        assert np.allclose(cl.fov.cpu().numpy()[0], [xe, ye], rtol=0, atol=1e-9)
        err += np.hypot(xe - p[8], ye - p[9])
    assert abs(float(cl.err_sum[0]) - err) <= 1e-9 * err


def _sampled_parity(oracle_mod, sc, p, x0, sol, st, n_sample, seed, obs=None):
    """Solve a seeded random sample of a full-size batch with the oracle and compare it with the CUDA path: statuses,
    and u0* / f* / active set wherever both converged (the north star's tolerances)."""
    lbx, ubx, lbg, ubg = sc.bounds()
    idx = np.sort(np.random.default_rng(seed).choice(p.shape[0], size=min(n_sample, p.shape[0]), replace=False))
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    ref = oracle_mod.solve(sp, sc.obstacle_table() if obs is None else obs[idx], p[idx], x0[idx], lbx, ubx, lbg, ubg,
                           obs_per_instance=obs is not None)
    tonp = lambda a: a[torch.as_tensor(idx, device=a.device)].cpu().numpy() if torch.is_tensor(a) else a[idx]
    sg, ig = tonp(st["return_status"]), tonp(st["iter_count"])
    agree = (ref["status"] == sg).mean()
    conv_agree = ((ref["status"] == 0) == (sg == 0)).mean()
    assert agree >= 0.97 and conv_agree >= 0.985, (agree, conv_agree, np.bincount(ref["status"], minlength=7), np.bincount(sg, minlength=7))
    both = (ref["status"] == 0) & (sg == 0)
    got = {k: tonp(sol[k])[both] for k in ("x", "f", "g")}
    r2 = {k: ref[k][both] for k in ("x", "f", "g")}; r2["status"] = ref["status"][both]
    _compare(r2, got, sg[both], ig[both], (lbx, ubx, lbg, ubg))
    assert (ref["iters"][both] == ig[both]).mean() >= 0.9
    return ref, idx


def test_full_size_properties(pkg, oracle_mod):
    """BASELINE config 2 size (4096 NMPC_TT instances): a 2048-instance sample against the oracle, and size-independent
    properties of every converged solve."""
    sc = pkg.SCENARIOS["nmpc_tt"]
    B = 4096
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=1000 * 2)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    sol = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    st = s.stats()
    ok = st["success"]
    ref, idx = _sampled_parity(oracle_mod, sc, p, x0, sol, st, 2048, seed=11)
    assert abs(ok[idx].mean() - (ref["status"] == 0).mean()) <= 0.01             # converged fraction = the oracle's
    assert np.all(st["iter_count"] <= 100) and np.all(st["iter_count"][ok] > 0)
    x, g = sol["x"][ok], sol["g"][ok]
    assert np.all(x >= lbx - 1e-12) and np.all(x <= ubx + 1e-12)                 # honour_original_bounds
    assert np.all(g <= ubg + 1e-4) and np.all(g >= lbg - 1e-4)                   # constr_viol_tol
    # f and g returned by the solve are the function values at the returned x (re-evaluated by nmpc_eval)
    ev = s.evaluate(sol["x"], p)
    assert np.allclose(ev["f"].cpu().numpy(), sol["f"], rtol=1e-13, atol=0)
    assert np.allclose(ev["g"].cpu().numpy(), sol["g"], rtol=0, atol=1e-10)
    # stationarity of the Lagrangian at converged points: grad f + J^T lam_g + lam_x ~ 0
    ev = s.evaluate(sol["x"], p, lam=sol["lam_g"])
    r = ev["grad"].cpu().numpy() + ev["jtv"].cpu().numpy() + sol["lam_x"]
    mult = np.maximum(1.0, np.maximum(np.abs(sol["lam_g"]).max(axis=1), np.abs(sol["lam_x"]).max(axis=1)))
    # x is clipped to the original bounds on return (honor_original_bounds): up to 1e-8 |b| off the internal iterate
    # (1.4e-7 on an active v = 14 bound), times the curvature of the T = 1 problem -> scale the tolerance with the
    # size of the terms that cancel
    gscale = np.maximum(np.abs(ev["grad"].cpu().numpy()).max(axis=1), mult)
    assert np.all(np.abs(r[ok]).max(axis=1) <= 1e-6 + 2e-6 * gscale[ok])
    # idempotence: re-solving from the solution converges to the same point
    sol2 = s(x0=sol["x"][ok][:256], p=p[ok][:256], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    ok2 = s.stats()["success"]
    assert ok2.mean() > 0.9
    # (the NLP is non-convex: a restart at mu = 0.1 may leave the basin, so this holds for most, not all, instances)
    same = np.abs(sol2["f"][ok2] - sol["f"][ok][:256][ok2]) <= 1e-6 * np.abs(sol["f"][ok][:256][ok2])
    assert same.mean() >= 0.8


@pytest.mark.parametrize("name,N,B,jitter", [("t_trajectory", 15, 65536, 0),          # BASELINE config 3 (one GPU's batch)
                                              ("plus_trajectory", 15, 65536, 0),
                                              ("10_obstacles", 15, 32768, 3),         # config 4: 262144 / 8 GPUs, 3 real obstacles jittered
                                              ("race_track_2", 30, 131072, 10)])      # config 5: 2x horizon, 1M / 8 GPUs, 10 obstacles jittered
def test_full_size_other_configs(pkg, oracle_mod, name, N, B, jitter):
    """BASELINE configs 3-5 at the per-GPU batch size, all on the device: a seeded sample against the oracle (2048
    instances; 512 at the doubled horizon, where a dense oracle solve costs ~0.3 s), and size-independent properties."""
    sc = pkg.SCENARIOS[name]
    if N != sc.N:
        sc = sc.with_horizon(N)
    dev = "cuda:0"
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=1000 * 3 + N + jitter)
    obs = None
    if jitter:
        rng = np.random.default_rng(77)
        obs = np.tile(sc.obstacle_table(), (B, 1, 1))
        obs[:, :jitter, :2] += rng.uniform(-100, 100, (B, jitter, 2))
        # reject layouts that put an obstacle within r + 20 of the start (move it away instead)
        d = np.linalg.norm(obs[:, :, :2] - p[:, None, :2], axis=2)
        close = d < obs[:, :, 2] + 20.0
        obs[:, :, 0] += np.where(close, 400.0, 0.0)
    T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    x0 = T(np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N)))
    pt = T(p); obt = None if obs is None else T(obs)
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    sol = s(x0=x0, p=pt, lbx=T(lbx), ubx=T(ubx), lbg=T(lbg), ubg=T(ubg), obstacles=obt)
    st = s.stats()
    ok = st["success"]
    x0n = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    ref, idx = _sampled_parity(oracle_mod, sc, p, x0n, sol, st, 512 if N == 30 else 2048, seed=13, obs=obs)
    assert abs(float(ok[torch.as_tensor(idx, device=dev)].double().mean()) - (ref["status"] == 0).mean()) <= 0.01
    assert float(ok.double().mean()) > 0.9
    assert int(st["iter_count"].max()) <= 100 and int(st["iter_count"][ok].min()) > 0
    x, g = sol["x"][ok], sol["g"][ok]
    assert bool((x >= T(lbx) - 1e-12).all()) and bool((x <= T(ubx) + 1e-12).all())
    assert bool((g <= T(ubg) + 1e-4).all()) and bool((g >= T(lbg) - 1e-4).all())
    ev = s.evaluate(sol["x"], pt, lam=sol["lam_g"], obstacles=obs)
    assert torch.allclose(ev["f"], sol["f"], rtol=1e-13, atol=0)
    assert torch.allclose(ev["g"], sol["g"], rtol=0, atol=1e-9)
    r = (ev["grad"] + ev["jtv"] + sol["lam_x"]).abs().amax(dim=1)
    mult = torch.maximum(sol["lam_g"].abs().amax(dim=1), sol["lam_x"].abs().amax(dim=1)).clamp(min=1.0)
    gscale = torch.maximum(ev["grad"].abs().amax(dim=1), mult)
    # The returned x is clipped to the original bounds (IPOPT's honor_original_bounds): up to 1e-8 off the internal
    # iterate.  Where the UAV passes almost exactly over the target the sqrt term has curvature 1/distance ~ 1e4
    # (SURVEY 7.3-5), so on a fraction of a percent of the instances that clip shows up as 1e-4 in the gradient of
    # an active control (tools/stat_probe.py).  Bound the bulk tightly and the outliers loosely.
    ratio = r[ok] / (1e-6 + 2e-6 * gscale[ok])
    assert float((ratio <= 1.0).double().mean()) >= 0.995, float((ratio <= 1.0).double().mean())
    assert float((r[ok] / (1.0 + gscale[ok])).max()) <= 1e-3
    # complementarity and sign of the bound multipliers (CasADi convention: lam > 0 at an upper bound)
    lam_g = sol["lam_g"][ok]
    ub, lb = T(ubg), T(lbg)
    fu, fl = torch.isfinite(ub), torch.isfinite(lb)
    zero = torch.zeros((), dtype=torch.float64, device=dev)
    up, lo = lam_g.clamp(min=0), (-lam_g).clamp(min=0)
    # IPOPT's unscaled complementarity tolerance compl_inf_tol = 1e-4
    assert float((up * torch.where(fu, ub - g, zero)).amax()) <= 1e-4 and float((lo * torch.where(fl, g - lb, zero)).amax()) <= 1e-4
    assert float(up[:, ~fu].amax() if (~fu).any() else 0.0) <= 1e-8 and float(lo[:, ~fl].amax() if (~fl).any() else 0.0) <= 1e-8


@pytest.mark.parametrize("name,N,B,steps", [("nmpc_tt", 15, 96, 40), ("race_track_2", 30, 24, 40)])
def test_closed_loop_teacher_forced_batch(pkg, oracle_mod, name, N, B, steps):
    """Teacher-forced closed loop on the scenarios where free-running loops separate (T = 1; the doubled horizon): the
    GPU loop runs, and at every step the oracle solves the GPU's own (p, warm start).  Per step: the statuses agree
    on (nearly) every instance, and wherever both converged u0*, f* and the active set agree to the north star's
    tolerances.  Infeasible NLPs that the loop itself produces are part of the population."""
    from mpc_implementation_b200.closed_loop import ClosedLoop
    sc = pkg.SCENARIOS[name]
    if N != sc.N:
        sc = sc.with_horizon(N)
    lbx, ubx, lbg, ubg = sc.bounds()
    p0, vw = pkg.random_instances(sc, B, seed=4242 + N)
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    cl = ClosedLoop(s, sc, p0, target_vw=vw)
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    agree = tot = conv_dis = n_both = 0
    acc_g, acc_r = {q: [] for q in ("x", "f", "g")}, {q: [] for q in ("x", "f", "g")}
    for k in range(steps):
        p = cl.p.cpu().numpy().copy(); w0 = cl.u_warm.cpu().numpy().copy()
        sol = cl.step(want_g=True)
        st = s.stats(); sg = st["return_status"].cpu().numpy(); ig = st["iter_count"].cpu().numpy()
        ref = oracle_mod.solve(sp, sc.obstacle_table(), p, w0, lbx, ubx, lbg, ubg, want_lam=False)
        agree += int((ref["status"] == sg).sum()); tot += B
        conv_dis += int(((ref["status"] == 0) != (sg == 0)).sum())
        both = (ref["status"] == 0) & (sg == 0)
        if k == 0:
            continue            # cold first step (u0 = 0): most T = 1 instances run into max_iter on both sides
        n_both += int(both.sum())
        for q in ("x", "f", "g"):
            acc_g[q].append(sol[q].cpu().numpy()[both]); acc_r[q].append(ref[q][both])
    cat = lambda d: {q: np.concatenate(v) for q, v in d.items()}
    _compare_bulk(cat(acc_r), cat(acc_g), (lbx, ubx, lbg, ubg))
    # converged / not converged is the decision the caller sees first: it agrees on >= 99 % of all solves; WHICH failure
    # exit an infeasible NLP takes (max_iter, Restoration_Failed, Infeasible_Problem_Detected) is decided late in long,
    # ill-conditioned runs and agrees a little less often (>= 95 %)
    assert conv_dis <= 0.01 * tot, (conv_dis, tot)
    assert agree >= 0.95 * tot, (agree, tot)
    assert n_both >= 0.8 * B * (steps - 1)


def test_bench_population_status_census(pkg, oracle_mod):
    """The bench workload itself (bench.py config 2: NMPC_TT, randomised states, closed loop): after five warm steps the
    oracle solves the same (p, warm start) population.  Converged / not-converged agrees on >= 99 % of the instances,
    the full status on >= 97 %, and the converged solutions agree to tolerance."""
    from mpc_implementation_b200.closed_loop import ClosedLoop
    sc = pkg.SCENARIOS["nmpc_tt"]
    B = 2048
    lbx, ubx, lbg, ubg = sc.bounds()
    p0, vw = pkg.random_instances(sc, B, seed=2000)
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    cl = ClosedLoop(s, sc, p0, target_vw=vw)
    for _ in range(5):
        cl.step()
    p = cl.p.cpu().numpy().copy(); w0 = cl.u_warm.cpu().numpy().copy()
    sol = cl.step(want_g=True)
    st = s.stats(); sg = st["return_status"].cpu().numpy(); ig = st["iter_count"].cpu().numpy()
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    ref = oracle_mod.solve(sp, sc.obstacle_table(), p, w0, lbx, ubx, lbg, ubg, want_lam=False)
    conf = np.zeros((7, 7), dtype=int)
    np.add.at(conf, (ref["status"], sg), 1)
    assert ((ref["status"] == 0) == (sg == 0)).mean() >= 0.99, conf
    assert np.trace(conf) >= 0.97 * B, conf
    assert conf[0, 0] >= 0.85 * B, conf
    both = (ref["status"] == 0) & (sg == 0)
    got = {q: sol[q].cpu().numpy()[both] for q in ("x", "f", "g")}
    r2 = {q: ref[q][both] for q in ("x", "f", "g")}
    _compare_bulk(r2, got, (lbx, ubx, lbg, ubg))
    # restoration / watchdog machinery is exercised by this population on both sides
    wc = s.work_counters()
    assert wc["resto_calls"] > 0 and ref["stats"][:, 2].sum() > 0


def test_restoration_outcomes(pkg, oracle_mod):
    """NLPs that are infeasible by construction (stage-0 rows violated: the UAV starts inside an obstacle's keep-out
    disc, or above the altitude ceiling by more than the bound relaxation): IPOPT's answer is the restoration phase
    and "Infeasible_Problem_Detected" (or Restoration_Failed / max_iter), never success; both sides agree, and the
    iterate handed back reduces the infeasibility (that is what the reference's loop would apply)."""
    sc = pkg.SCENARIOS["nmpc_tt"]
    lbx, ubx, lbg, ubg = sc.bounds()
    B = 32
    p, _ = pkg.random_instances(sc, B, seed=99)
    ob = sc.obstacle_table()
    p[:16, 0] = ob[1, 0] + 10.0; p[:16, 1] = ob[1, 1] - 5.0          # inside obstacle 2 (r_uav + r_obs = 35)
    p[16:, 2] = 150.0 + 10.0 ** np.linspace(-5.5, -2, 16)            # above z <= 150 (relaxed: 150 + 1.5e-6)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    ref = oracle_mod.solve(sp, ob, p, x0, lbx, ubx, lbg, ubg)
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    sol = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    sg = s.stats()["return_status"]
    assert (ref["status"] != 0).all() and (sg != 0).all()
    assert np.isin(sg, (1, 2, 6)).all() and (sg == 6).sum() >= B // 2, sg
    assert (ref["status"] == sg).mean() >= 0.8, (ref["status"], sg)
    assert s.work_counters()["resto_calls"] >= B // 2
    assert np.isfinite(sol["x"]).all() and np.all(sol["x"] >= lbx - 1e-12) and np.all(sol["x"] <= ubx + 1e-12)
    # independent certificate of "Infeasible_Problem_Detected" (as tests/test_oracle_solve.py::
    # test_infeasible_exit_is_a_local_minimiser_of_the_violation): an LP on the linearised rows around the point the
    # KERNEL returned finds no direction that decreases the l1 violation
    from scipy.optimize import linprog
    n, m = sc.n_w, sc.n_g
    fu, fl = np.isfinite(ubg), np.isfinite(lbg)
    for i in np.flatnonzero(sg == 6)[:12]:
        g = sol["g"][i]
        J = oracle_mod.evaluate(sp, ob, sol["x"][i], p[i])["J"]
        theta = np.maximum(lbg - g, 0).sum() + np.maximum(g - ubg, 0).sum()
        assert theta > 1e-6, (i, theta)
        A = np.hstack([J, np.eye(m), -np.eye(m)])
        res = linprog(np.concatenate([np.zeros(n), np.ones(2 * m)]),
                      A_ub=np.vstack([A[fu], -A[fl]]), b_ub=np.concatenate([(ubg - g)[fu], (g - lbg)[fl]]),
                      bounds=[(max(lbx[k] - sol["x"][i][k], -1e-2), min(ubx[k] - sol["x"][i][k], 1e-2)) for k in range(n)] + [(0, None)] * (2 * m),
                      method="highs")
        assert res.status == 0 and theta - res.fun <= 1e-5 * theta + 1e-8, (i, theta, res.fun)


def test_schedule_independence(pkg):
    """The fetch order of the persistent kernel (nmpc_set_order / the library's longest-first order) and the lockstep
    alignment of the warps of a block are pure scheduling: results are bit-identical whatever the order."""
    sc = pkg.SCENARIOS["nmpc_tt"]
    B = 1500                                           # more instances than resident warps -> the queue is used
    dev = "cuda:0"
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=77)
    T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    x0 = T(np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))); pt = T(p)
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=B)
    args = dict(x0=x0, p=pt, lbx=T(lbx), ubx=T(ubx), lbg=T(lbg), ubg=T(ubg))
    a = s(**args); sa = {k: v.clone() for k, v in s.stats().items()}                 # first call: natural order
    b = s(**args); sb = {k: v.clone() for k, v in s.stats().items()}                 # second call: longest-first (auto)
    perm = torch.randperm(B, device=dev)
    c = s(order=perm, **args); sc_ = {k: v.clone() for k, v in s.stats().items()}     # explicit random order
    for other, so in ((b, sb), (c, sc_)):
        assert torch.equal(sa["return_status"], so["return_status"]) and torch.equal(sa["iter_count"], so["iter_count"])
        for k in ("x", "f", "g", "lam_x", "lam_g"):
            assert torch.equal(a[k], other[k]), k


def test_pipelined_closed_loop_is_the_same_loop(pkg):
    """PipelinedClosedLoop (sub-batches on their own handles and streams) advances every instance exactly like the
    single-batch ClosedLoop: bit-identical states, warm starts, error sums and solver stats after several steps."""
    from mpc_implementation_b200.closed_loop import ClosedLoop, PipelinedClosedLoop
    sc = pkg.SCENARIOS["nmpc_tt"]
    B = 700
    p, vw = pkg.random_instances(sc, B, seed=123)
    mk = lambda n: pkg.nlpsol("solver", "ipm", sc, max_batch=n)
    one = ClosedLoop(mk(B), sc, p, target_vw=vw)
    many = PipelinedClosedLoop(mk, sc, p, target_vw=vw, pipelines=3)
    for _ in range(4):
        one.step(); many.step()
    torch.cuda.synchronize()
    assert torch.equal(one.p, many.p) and torch.equal(one.u_warm, many.u_warm) and torch.equal(one.err_sum, many.err_sum)
    so, sm = one.solver.stats(), many.stats()
    assert torch.equal(so["return_status"], sm["return_status"]) and torch.equal(so["iter_count"], sm["iter_count"])


def test_handle_reuse_across_batch_sizes(pkg):
    """One handle, calls of changing batch size (the launch of call n prepares the queue, counters and fetch order of
    call n+1): every call returns what a fresh solver returns for the same inputs."""
    sc = pkg.SCENARIOS["nmpc_tt"]
    dev = "cuda:0"
    lbx, ubx, lbg, ubg = sc.bounds()
    T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
    p, _ = pkg.random_instances(sc, 1300, seed=5)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (1300, sc.N))
    bt = (T(lbx), T(ubx), T(lbg), T(ubg))
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=64)
    for B in (64, 1300, 1300, 200, 1300, 1, 200, 200):
        got = s(x0=T(x0[:B]), p=T(p[:B]), lbx=bt[0], ubx=bt[1], lbg=bt[2], ubg=bt[3], want_g=False, want_lam=False)
        gst = {k: v.clone() for k, v in s.stats().items()}
        f = pkg.nlpsol("fresh", "ipm", sc, max_batch=B)
        ref = f(x0=T(x0[:B]), p=T(p[:B]), lbx=bt[0], ubx=bt[1], lbg=bt[2], ubg=bt[3], want_g=False, want_lam=False)
        rst = f.stats()
        assert torch.equal(got["x"], ref["x"]) and torch.equal(got["f"], ref["f"]), B
        assert torch.equal(gst["return_status"], rst["return_status"]) and torch.equal(gst["iter_count"], rst["iter_count"]), B
        wc = s.work_counters(); wf = f.work_counters()
        assert wc["factorizations"] == wf["factorizations"] and wc["ls_trials"] == wf["ls_trials"], B


def test_async_host_entry_point(pkg):
    """nmpc_solve_host_async + nmpc_synchronize (solver(..., blocking=False); solver.wait()) returns what the blocking
    host call returns; two handles driven this way can be in flight together."""
    sc = pkg.SCENARIOS["t_trajectory"]
    B = 64
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, 2 * B, seed=3)
    pin = lambda a: (lambda t: (t.copy_(torch.from_numpy(np.ascontiguousarray(a))), t.numpy())[1])(torch.empty(a.shape, dtype=torch.float64).pin_memory())
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (2 * B, sc.N))
    s1, s2 = pkg.nlpsol("a", "ipm", sc, max_batch=B), pkg.nlpsol("b", "ipm", sc, max_batch=B)
    ref1 = s1(x0=x0[:B], p=p[:B], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    ref2 = s2(x0=x0[B:], p=p[B:], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    st1 = {k: v.copy() for k, v in s1.stats().items()}
    pa, pb, xa, xb = pin(p[:B]), pin(p[B:]), pin(x0[:B]), pin(x0[B:])
    o1 = s1(x0=xa, p=pa, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, blocking=False)       # both enqueued before either is waited for
    o2 = s2(x0=xb, p=pb, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, blocking=False)
    s1.wait(); s2.wait()
    for k in ("x", "f", "g", "lam_x", "lam_g"):
        assert np.array_equal(o1[k], ref1[k]) and np.array_equal(o2[k], ref2[k]), k
    assert np.array_equal(s1.stats()["return_status"], st1["return_status"]) and np.array_equal(s1.stats()["iter_count"], st1["iter_count"])
    assert s1.stats()["success"].dtype == bool


def test_edge_cases(pkg):
    sc = pkg.SCENARIOS["t_trajectory"]
    lbx, ubx, lbg, ubg = sc.bounds()
    s = pkg.nlpsol("solver", "ipm", sc, max_batch=4)
    out = s(x0=np.zeros((0, sc.n_w)), p=np.zeros((0, 11)), lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)       # empty batch
    assert out["x"].shape == (0, sc.n_w)
    with pytest.raises(ValueError):
        s(x0=np.zeros((2, sc.n_w)), p=np.zeros((3, 11)), lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)          # ragged
    with pytest.raises(ValueError):
        s(x0=np.zeros(sc.n_w), p=np.zeros(11), lbx=lbx, ubx=ubx, lbg=lbg)                              # missing bound
    # UAV exactly above the target: sqrt is not differentiable there -> reported, not hidden
    p = np.array(list(sc.x_init) + [99.0, 150.0, 0.0])
    sol = s(x0=np.zeros(sc.n_w), p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    assert int(s.stats()["return_status"][0]) in (0, 1, 2, 3, 4, 5, 6)
    # growing past max_batch re-creates the handle transparently
    pb, _ = pkg.random_instances(sc, 9, 1)
    out = s(x0=np.zeros((9, sc.n_w)), p=pb, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    assert out["x"].shape == (9, sc.n_w)


# ------------------------------------------------------------------------------------------------------------------
# NON-REFERENCE fast mode: multiplier / mu warm start.  Off by default; parity against the oracle's restatement of
# IPOPT's WarmStartIterateInitializer.
# ------------------------------------------------------------------------------------------------------------------
WARM_OPTS = {"ipopt": {"max_iter": 100, "warm_start_init_point": "yes", "mu_init": 1e-4}}


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["t_trajectory", "race_track_2", "gimbal_less"])
def test_warm_start_parity(pkg, oracle_mod, name):
    """Re-solve from the primal-dual solution: the kernel follows the oracle's warm-started iterates (same statuses and iteration
    counts, same solutions), takes far fewer iterations than the cold start, and a NaN marker cold-starts an instance bit for
    bit like a call without guesses.  Without warm_start_init_point the guesses are ignored, as in IPOPT."""
    sc = pkg.SCENARIOS[name]
    B = 96
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=31)
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov, sc.model)
    obs = sc.obstacle_table()
    r0 = oracle_mod.solve(sp, obs, p, np.zeros((B, sc.n_w)), lbx, ubx, lbg, ubg)
    lx = r0["lam_x"].copy(); lx[::4, 0] = np.nan
    oracle_mod.set_option("ws_mu_init", 1e-4)
    try:
        ref = oracle_mod.solve(sp, obs, p, r0["x"], lbx, ubx, lbg, ubg, lam_x0=lx, lam_g0=r0["lam_g"])
    finally:
        oracle_mod.clear_options()
    s = pkg.nlpsol("w", "ipm", sc, WARM_OPTS, max_batch=B)
    sol = s(x0=r0["x"], p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, lam_x0=lx, lam_g0=r0["lam_g"])
    st = s.stats()
    assert (st["return_status"] == ref["status"]).mean() >= 0.97
    both = (st["return_status"] == 0) & (ref["status"] == 0)
    assert both.mean() > 0.85
    assert (st["iter_count"][both] == ref["iters"][both]).mean() >= 0.95
    rf = np.abs(sol["f"].ravel() - ref["f"])[both] / np.maximum(1.0, np.abs(ref["f"][both]))
    assert rf.max() <= F_RTOL
    ru = np.abs(sol["x"][both, :sc.nu] - ref["x"][both, :sc.nu]).max(axis=1) / np.abs(ref["x"][both, :sc.nu]).max(axis=1)
    assert np.mean(ru <= U0_RTOL) >= 0.97 and ru.max() <= 1e-3
    # cold reference on the same starts
    c = pkg.nlpsol("c", "ipm", sc, max_batch=B)
    cold = c(x0=r0["x"], p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, lam_x0=lx, lam_g0=r0["lam_g"])      # guesses ignored: not enabled
    cst = c.stats()
    cold2 = c(x0=r0["x"], p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    assert np.array_equal(cold["x"], cold2["x"]) and np.array_equal(cst["iter_count"], c.stats()["iter_count"])
    assert np.array_equal(sol["x"][::4], cold["x"][::4]) and np.array_equal(st["iter_count"][::4], cst["iter_count"][::4])
    w = both.copy(); w[::4] = False
    assert st["iter_count"][w].mean() < 0.7 * cst["iter_count"][w].mean()


@pytest.mark.gpu
def test_warm_duals_closed_loop(pkg, oracle_mod):
    """The fused epilogue's dual shift: ClosedLoop(warm_duals=True) against the oracle driven with the same shift on the host."""
    import torch
    from mpc_implementation_b200.closed_loop import ClosedLoop
    sc = pkg.SCENARIOS["t_trajectory"]
    B, K = 64, 6
    lbx, ubx, lbg, ubg = sc.bounds()
    p0, vw = pkg.random_instances(sc, B, seed=41)
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    obs = sc.obstacle_table()
    cl = ClosedLoop(pkg.nlpsol("w", "ipm", sc, WARM_OPTS, max_batch=B), sc, p0, target_vw=vw, warm_duals=True)
    R = 5 + sc.n_obs
    p = p0.copy(); u = np.zeros((B, sc.n_w)); lx = np.zeros((B, sc.n_w)); lg = np.zeros((B, sc.n_g)); lx[:, 0] = np.nan
    oracle_mod.set_option("ws_mu_init", 1e-4)
    try:
        for k in range(K):
            cl.step()
            stg = cl.solver.stats()
            r = oracle_mod.solve(sp, obs, p, u, lbx, ubx, lbg, ubg, lam_x0=lx, lam_g0=lg)
            gs, gi = stg["return_status"].cpu().numpy(), stg["iter_count"].cpu().numpy()
            assert (gs == r["status"]).mean() >= 0.95, k
            same = (gs == 0) & (r["status"] == 0)
            assert (gi[same] == r["iters"][same]).mean() >= 0.9, k
            # teacher forcing: the oracle continues from the GPU's iterate so that single flips do not compound
            x = cl.last["x"].cpu().numpy()
            assert np.abs(x[same] - r["x"][same]).max() <= 1e-4
            ok = gs == 0
            lx = cl.lam_x0.cpu().numpy().copy(); lg = cl.lam_g0.cpu().numpy().copy()
            assert np.array_equal(np.isnan(lx[:, 0]), ~ok)                     # marker after a failed solve
            # the shifted multipliers are this solve's own (compare with the oracle's where both converged)
            ex = np.concatenate([r["lam_x"].reshape(B, sc.N, 6)[:, 1:], r["lam_x"].reshape(B, sc.N, 6)[:, -1:]], 1).reshape(B, -1)
            eg = np.concatenate([r["lam_g"].reshape(B, sc.N + 1, R)[:, 1:], r["lam_g"].reshape(B, sc.N + 1, R)[:, -1:]], 1).reshape(B, -1)
            assert np.abs(lx[same][:, 1:] - ex[same][:, 1:]).max() <= 1e-5 * max(1.0, np.abs(ex[same]).max())
            assert np.abs(lg[same] - eg[same]).max() <= 1e-5 * max(1.0, np.abs(eg[same]).max())
            p = cl.p.cpu().numpy().copy(); u = cl.u_warm.cpu().numpy().copy()
    finally:
        oracle_mod.clear_options()


# ------------------------------------------------------------------------------------------------------------------
# free-running closed loop (nmpc_run_closed_loop): K steps of every instance in one launch, no batch-wide barrier between steps
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,N,B,K,phases", [("nmpc_tt", 15, 300, 7, False), ("t_trajectory", 15, 257, 9, True), ("race_track_2", 30, 96, 4, False),
                                               ("10_obstacles", 15, 64, 5, False), ("gimbal_less", 15, 200, 6, False)])
def test_free_running_loop_is_the_same_loop(pkg, name, N, B, K, phases):
    """Bit-identical to K consecutive nmpc_solve_and_step calls: states, warm starts, error sums, and the status / iteration count
    of every solve -- with a constant per-instance target input and with the device schedule (+ per-instance phases)."""
    import torch
    from mpc_implementation_b200.closed_loop import ClosedLoop
    sc = pkg.SCENARIOS[name]
    if N != sc.N:
        sc = sc.with_horizon(N)
    p0, vw = pkg.random_instances(sc, B, seed=17)
    kw = dict(phase=np.random.default_rng(3).integers(0, 400, B)) if phases else dict(target_vw=vw)
    a = ClosedLoop(pkg.nlpsol("a", "ipm", sc, max_batch=B), sc, p0, **kw)
    b = ClosedLoop(pkg.nlpsol("b", "ipm", sc, max_batch=B), sc, p0, **kw)
    st, it = [], []
    for k in range(K):
        a.step(); s = a.solver.stats()
        st.append(s["return_status"].clone()); it.append(s["iter_count"].clone())
    out = b.run_free(K)
    torch.cuda.synchronize()
    assert torch.equal(a.p, b.p) and torch.equal(a.u_warm, b.u_warm) and torch.equal(a.err_sum, b.err_sum) and torch.equal(a.fov, b.fov)
    assert torch.equal(torch.stack(st), out["status_log"]) and torch.equal(torch.stack(it), out["iters_log"])
    assert torch.equal(out["converged"], (torch.stack(st) == 0).sum(dim=0).to(torch.int32))
    # and the two loops stay interchangeable afterwards (schedule counter, fetch order)
    a.step(); b.step(); torch.cuda.synchronize()
    assert torch.equal(a.p, b.p) and torch.equal(a.u_warm, b.u_warm)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["nmpc_tt", "race_track_2", "gimbal_less"])
def test_lam_p(pkg, name):
    """sol['lam_p'] (CasADi returns it next to lam_x / lam_g, NMPC_TT.py:358-365): minus the gradient of f + lam_g^T g with respect
    to p at the returned point, against autograd of the literal NLP -- and the same sign convention as lam_x."""
    from oracle import nlp_ref
    sc = pkg.SCENARIOS[name]
    B = 6
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = pkg.random_instances(sc, B, seed=9)
    s = pkg.nlpsol("s", "ipm", sc, max_batch=B)
    sol = s(x0=np.zeros((B, sc.n_w)), p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    assert sol["lam_p"].shape == (B, sc.n_p)
    for b in range(B):
        pt = torch.tensor(p[b], requires_grad=True); wt = torch.tensor(sol["x"][b], requires_grad=True); lt = torch.tensor(sol["lam_g"][b])
        if sc.model:
            rs = nlp_ref.RefSpec5(sc.T, sc.N)
            Lg = nlp_ref.objective5(rs, wt, pt) + (lt * nlp_ref.constraints5(rs, wt, pt)).sum()
        else:
            rs = nlp_ref.RefSpec(T=sc.T, N=sc.N, obstacles=sc.obstacles, uav_r=sc.uav_r, w1=sc.w1, w2=sc.w2)
            Lg = nlp_ref.objective(rs, wt, pt) + (lt * nlp_ref.constraints(rs, wt, pt)).sum()
        gp, gw = torch.autograd.grad(Lg, [pt, wt])
        scale = max(1.0, float(gp.abs().max()))
        assert np.abs(sol["lam_p"][b] + gp.numpy()).max() <= 1e-9 * scale, (sol["lam_p"][b], gp)
        if s.stats()["return_status"][b] == 0:      # same convention as lam_x: grad_x L + lam_x = 0 at a solution
            assert np.abs(sol["lam_x"][b] + gw.numpy()).max() <= 1e-5 * max(1.0, float(gw.abs().max()))
