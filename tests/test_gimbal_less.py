"""Model variant of SURVEY 8f-4: the gimbal-less 5-state / 3-control tracker of MATLAB/Dynamic Obstacles/NMPC_TT.m:25-37
(distance-only cost :100-104, rows [z, theta] per stage :107-111, closed loop :138-170 with shift1.m).

CPU: the oracle's model-1 functions against a literal autograd restatement of the MATLAB script (oracle/nlp_ref.py: *5),
KKT certificates of the committed fixture, the C ABI's size functions.  GPU (-m gpu): the CUDA path -- which runs the model on
the 8-state machinery with absent camera controls -- against the oracle, which solves the true 3N-variable NLP."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle import nlp_ref

GOLD = Path(__file__).resolve().parent / "golden"
U0_RTOL, F_RTOL = 1e-6, 1e-8      # the north star's tolerances


def _setup(pkg, oracle_mod):
    sc = pkg.SCENARIOS["gimbal_less"]
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, model=1)
    return sc, sp, np.zeros((0, 3)), sc.bounds()


def test_scenario_matches_the_script(pkg):
    sc = pkg.SCENARIOS["gimbal_less"]
    assert (sc.T, sc.N, sc.n_obs, sc.steps, sc.model) == (0.2, 15, 0, 100, 1)            # NMPC_TT.m:9-10, :144
    assert (sc.n_w, sc.n_g, sc.n_p) == (45, 32, 8)                                        # :37, :107-111, :114
    lbx, ubx, lbg, ubg = sc.bounds()
    a, b, c, d = nlp_ref.bounds5(nlp_ref.RefSpec5(sc.T, sc.N))
    assert np.array_equal(lbx, a) and np.array_equal(ubx, b) and np.array_equal(lbg, c) and np.array_equal(ubg, d)
    assert sc.x_init == (90.0, 150.0, 80.0, 0.0, 0.0) and sc.target_init == (100.0, 150.0, 0.0)   # :138-139
    assert sc.schedule(0) == (15.0, 0.12) and sc.schedule(99) == (15.0, 0.12)            # shift1.m:9


def test_oracle_functions_match_autograd(pkg, oracle_mod):
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod)
    rs = nlp_ref.RefSpec5(sc.T, sc.N)
    rng = np.random.default_rng(0)
    p, _ = pkg.random_instances(sc, 4, seed=1)
    for b in range(4):
        w = lbx + (ubx - lbx) * rng.uniform(size=sc.n_w); lam = rng.normal(size=sc.n_g)
        a = nlp_ref.eval_all5(rs, w, p[b], lam, 1.0)
        o = oracle_mod.evaluate(sp, obs, w, p[b], lam_g=lam, sigma=1.0, hessian=True)
        assert abs(a["f"] - o["f"]) <= 1e-12 * abs(a["f"])
        assert np.abs(a["g"] - o["g"]).max() <= 1e-12 and np.abs(a["grad"] - o["grad"]).max() <= 1e-11
        assert np.abs(a["J"] - o["J"]).max() <= 1e-12 and np.abs(a["H"] - o["H"]).max() <= 1e-11 * max(1.0, np.abs(a["H"]).max())


def test_oracle_fixture_and_kkt(pkg, oracle_mod):
    """The committed fixture is reproduced bit for bit, and its converged solutions satisfy the KKT conditions of the literal NLP."""
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod)
    G = np.load(GOLD / "solves_gimbal_less.npz")
    r = oracle_mod.solve(sp, obs, G["p"], G["x0"], lbx, ubx, lbg, ubg)
    assert np.array_equal(r["status"], G["status"]) and np.array_equal(r["iters"], G["iters"]) and np.array_equal(r["x"], G["x"])
    rs = nlp_ref.RefSpec5(sc.T, sc.N)
    idx = np.flatnonzero(G["status"] == 0)
    assert len(idx) >= 10
    for i in idx[:4]:
        x, lam_g, lam_x = G["x"][i], G["lam_g"][i], G["lam_x"][i]
        d = nlp_ref.eval_all5(rs, x, G["p"][i])
        assert abs(d["f"] - G["f"][i]) <= 1e-10 * abs(d["f"]) and np.abs(d["g"] - G["g"][i]).max() <= 1e-9
        mult = max(1.0, np.abs(lam_g).max(), np.abs(lam_x).max())
        assert np.abs(d["grad"] + d["J"].T @ lam_g + lam_x).max() <= 1e-6 + 1e-7 * mult
        assert np.all(d["g"] <= ubg + 1e-6) and np.all(d["g"] >= lbg - 1e-6) and np.all(x <= ubx + 1e-12) and np.all(x >= lbx - 1e-12)
        su, sl = ubg - d["g"], d["g"] - lbg
        assert np.all(lam_g[su > 1e-3] <= 1e-5) and np.all(lam_g[sl > 1e-3] >= -1e-5)
        assert np.all(lam_x[ubx - x > 1e-4] <= 1e-4) and np.all(lam_x[x - lbx > 1e-4] >= -1e-4)


def test_oracle_closed_loop_of_the_script(pkg, oracle_mod):
    """NMPC_TT.m:146-170: 100 steps from x0 = [90;150;80;0;0], xs = [100;150;0], u0 = 0.  The first two solves are the mirror-
    symmetric start (psi = 0, target dead ahead: a saddle, like the Python scripts' first step) and run into max_iter; every
    later one converges, and the UAV ends up circling over the target."""
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod)
    x0, xs, u0 = np.array(sc.x_init), np.array(sc.target_init), np.zeros((sc.N, 3))
    status = []
    for k in range(sc.steps):
        r = oracle_mod.solve(sp, obs, np.concatenate([x0, xs])[None], u0.reshape(1, -1), lbx, ubx, lbg, ubg, nthreads=1)
        x0, u0, xs = nlp_ref.shift1(sc.T, x0, r["x"].reshape(sc.N, 3), xs, sc.schedule(k))
        status.append(int(r["status"][0]))
    assert status[2:] == [0] * (sc.steps - 2)
    assert np.hypot(x0[0] - xs[0], x0[1] - xs[1]) < 40.0 and 75.0 <= x0[2] <= 150.0


def test_c_abi_sizes(pkg):
    L = pkg._ffi.lib()
    sp = pkg._ffi.NmpcSpec(0.2, 15, 0, 1.0, 0.0, 1.0, 1.0, 100, 1, 1e-8, 1, 1, 1)
    assert (L.nmpc_n_w(C.byref(sp)), L.nmpc_n_g(C.byref(sp)), L.nmpc_n_p(C.byref(sp))) == (45, 32, 8)
    sp.model = 0; sp.n_obs = 3
    assert (L.nmpc_n_w(C.byref(sp)), L.nmpc_n_g(C.byref(sp)), L.nmpc_n_p(C.byref(sp))) == (90, 128, 11)


# ------------------------------------------------------------------------------------------------------------------
def _np(d):
    import torch
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in d.items()}


@pytest.mark.gpu
def test_gpu_functions(pkg, oracle_mod):
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod)
    B = 8
    p, _ = pkg.random_instances(sc, B, seed=2)
    rng = np.random.default_rng(3)
    w = lbx + (ubx - lbx) * rng.uniform(size=(B, sc.n_w)); lam = rng.normal(size=(B, sc.n_g)); v = rng.normal(size=(B, sc.n_w))
    s = pkg.nlpsol("s", "ipm", sc, max_batch=B)
    e = _np(s.evaluate(w, p, lam=lam, v=v))
    for b in range(B):
        o = oracle_mod.evaluate(sp, obs, w[b], p[b], lam_g=lam[b], hessian=True)
        assert abs(o["f"] - float(e["f"][b])) <= 1e-12 * abs(o["f"])
        assert np.abs(o["g"] - e["g"][b]).max() <= 1e-12 and np.abs(o["grad"] - e["grad"][b]).max() <= 1e-11
        assert np.abs(o["J"].T @ lam[b] - e["jtv"][b]).max() <= 1e-11 and np.abs(o["H"] @ v[b] - e["hv"][b]).max() <= 1e-10


def _parity(ref, sol, st, bounds, nu):
    lbx, ubx, lbg, ubg = bounds
    assert (ref["status"] == st["return_status"]).mean() >= 0.98
    both = (ref["status"] == 0) & (st["return_status"] == 0)
    assert both.sum() >= 0.98 * (ref["status"] == 0).sum()
    rf = np.abs(ref["f"] - sol["f"].ravel())[both] / np.maximum(1.0, np.abs(ref["f"][both]))
    ru = np.abs(ref["x"][both, :nu] - sol["x"][both, :nu]).max(axis=1) / np.abs(ref["x"][both, :nu]).max(axis=1)
    assert rf.max() <= F_RTOL and ru.max() <= U0_RTOL, (rf.max(), ru.max())
    tol = 1e-6        # identical active set (controls and rows)
    for a, b_, lo, hi in ((ref["x"][both], sol["x"][both], lbx, ubx), (ref["g"][both], sol["g"][both], lbg, ubg)):
        assert np.array_equal(a >= hi - tol * np.maximum(1, np.abs(hi)), b_ >= hi - tol * np.maximum(1, np.abs(hi)))
        assert np.array_equal(a <= lo + tol * np.maximum(1, np.abs(lo)), b_ <= lo + tol * np.maximum(1, np.abs(lo)))
    assert (st["iter_count"][both] == ref["iters"][both]).mean() >= 0.95       # same iterates, not just the same answers


@pytest.mark.gpu
def test_gpu_golden_and_random_parity(pkg, oracle_mod):
    sc, sp, obs, bounds = _setup(pkg, oracle_mod)
    lbx, ubx, lbg, ubg = bounds
    G = np.load(GOLD / "solves_gimbal_less.npz")
    s = pkg.nlpsol("s", "ipm", sc, max_batch=len(G["f"]))
    sol = _np(s(x0=G["x0"], p=G["p"], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)); st = _np(s.stats())
    assert np.array_equal(G["status"], st["return_status"])
    _parity({k: G[k] for k in ("x", "f", "g", "status", "iters")}, sol, st, bounds, sc.nu)
    B = 768
    p, _ = pkg.random_instances(sc, B, seed=77)
    x0 = np.tile(np.array([16.0, 0.0, 0.0]), (B, sc.N)); x0[: B // 4] = 0.0
    ref = oracle_mod.solve(sp, obs, p, x0, lbx, ubx, lbg, ubg)
    s = pkg.nlpsol("s", "ipm", sc, max_batch=B)
    sol = _np(s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)); st = _np(s.stats())
    assert (ref["status"] == 0).mean() > 0.9
    _parity(ref, sol, st, bounds, sc.nu)


@pytest.mark.gpu
def test_gpu_closed_loop_of_the_script(pkg, oracle_mod):
    """The fused on-device loop (solve + shift1.m in the kernel epilogue) against the oracle driven by the literal shift1 on the
    host, all 100 steps of NMPC_TT.m; and the two-launch path (nmpc_solve + nmpc_step) bit for bit against the fused one."""
    import torch
    from mpc_implementation_b200.closed_loop import ClosedLoop
    sc, sp, obs, (lbx, ubx, lbg, ubg) = _setup(pkg, oracle_mod)
    p0 = np.array([list(sc.x_init) + list(sc.target_init)])
    cl = ClosedLoop(pkg.nlpsol("c", "ipm", sc, max_batch=1), sc, p0)
    s2 = pkg.nlpsol("u", "ipm", sc, max_batch=1)
    dev = cl.p.device
    p2 = torch.as_tensor(p0, device=dev).clone(); u2 = torch.zeros((1, sc.n_w), dtype=torch.float64, device=dev)
    vw = torch.tensor([[15.0, 0.12]], dtype=torch.float64, device=dev)
    x0, xs, u0 = np.array(sc.x_init), np.array(sc.target_init), np.zeros((sc.N, 3))
    for k in range(sc.steps):
        cl.step()
        sol = s2(x0=u2, p=p2, lbx=cl.lbx, ubx=cl.ubx, lbg=cl.lbg, ubg=cl.ubg, want_g=False, want_lam=False)
        s2.step(sol["x"], p2, u2, vw)
        r = oracle_mod.solve(sp, obs, np.concatenate([x0, xs])[None], u0.reshape(1, -1), lbx, ubx, lbg, ubg, nthreads=1)
        assert int(cl.solver.stats()["return_status"][0]) == int(r["status"][0]), k
        x0, u0, xs = nlp_ref.shift1(sc.T, x0, r["x"].reshape(sc.N, 3), xs, sc.schedule(k))
    torch.cuda.synchronize()
    assert torch.equal(cl.p, p2) and torch.equal(cl.u_warm, u2)
    pg = cl.p.cpu().numpy()[0]
    assert np.abs(pg - np.concatenate([x0, xs])).max() <= 1e-6 * np.abs(pg).max()
