"""Generates tests/golden/*.npz from the CPU oracle (oracle/nmpc_oracle.cpp).

The reference ships no golden vectors and CasADi/IPOPT is not installable here (parity unpinned), so
these fixtures pin (a) the oracle against regressions and (b) the CUDA path against the oracle at the
exact inputs of the reference scripts' first closed-loop steps plus seeded Monte-Carlo instances.
Every stored solution is KKT-certified independently in tests/test_oracle_solve.py.
Run:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import b200nmpc  # noqa: E402  (scenario registry only; no GPU needed)
import oracle    # noqa: E402
from oracle import nlp_ref  # noqa: E402

OUT = Path(__file__).resolve().parent
CASES = [("nmpc_tt", 6, 8), ("t_trajectory", 6, 8), ("plus_trajectory", 4, 4), ("race_trajectory_1", 4, 4),
         ("race_track_2", 6, 8), ("10_obstacles", 4, 8), ("gimbal_less", 6, 8)]


def closed_loop_cases(sc, steps):
    """(p, x0) of the first `steps` solves of the script's own loop, driven by the oracle."""
    sp = oracle.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov, sc.model)
    obs = sc.obstacle_table()
    lbx, ubx, lbg, ubg = sc.bounds()
    x0 = np.array(sc.x_init); xs = np.array(sc.target_init); u0 = np.zeros((sc.nu, sc.N))
    P, X0 = [], []
    for i in range(steps):
        p = np.concatenate([x0, xs]); w0 = u0.T.reshape(-1)
        P.append(p); X0.append(w0)
        r = oracle.solve(sp, obs, p, w0, lbx, ubx, lbg, ubg, nthreads=1)
        u = r["x"][0].reshape(sc.N, sc.nu).T
        if sc.model:      # MATLAB/Dynamic Obstacles/shift1.m (rows = stages)
            x0, u0, xs = nlp_ref.shift1(sc.T, x0, u.T, xs, sc.schedule(i)); u0 = u0.T
        else:
            x0, u0, xs = nlp_ref.shift_timestep(sc.T, x0, u, xs, sc.schedule(i))
    return np.array(P), np.array(X0)


def main():
    only = sys.argv[1:]      # optional: scenario names to (re)generate
    for name, steps, nrand in CASES:
        if only and name not in only:
            continue
        sc = b200nmpc.SCENARIOS[name]
        sp = oracle.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov, sc.model)
        obs = sc.obstacle_table()
        lbx, ubx, lbg, ubg = sc.bounds()
        P, X0 = closed_loop_cases(sc, steps)
        pr, _ = b200nmpc.random_instances(sc, nrand, seed=1234)
        # warm-ish starts for the random instances: feasible mid-range controls
        xr = np.tile(np.array([16.0, 0.0, 0.0, 0.0, 0.0, 0.0][:sc.nu]), (nrand, sc.N))
        P = np.concatenate([P, pr]); X0 = np.concatenate([X0, xr])
        r = oracle.solve(sp, obs, P, X0, lbx, ubx, lbg, ubg)
        np.savez_compressed(OUT / f"solves_{name}.npz", p=P, x0=X0, x=r["x"], f=r["f"], g=r["g"], lam_x=r["lam_x"],
                            lam_g=r["lam_g"], status=r["status"], iters=r["iters"], n_loop=steps)
        print(name, "status", np.bincount(r["status"], minlength=3), "iters", r["iters"])


if __name__ == "__main__":
    main()
