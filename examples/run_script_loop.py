"""The main loop of a reference script (Python/NMPC_TT.py:346-402 and siblings), written the way the script writes it --
one solver call per MPC step with host arrays, the shift on the host -- with the one swapped line:

    solver = ca.nlpsol('solver', 'ipopt', nlp_prob, opts)        # reference, :267
    solver = b200nmpc.nlpsol('solver', 'ipm', scenario, opts)    # here

and at the end the number the script prints: sum_i || FOVcentre_{i+1} - target_i ||  (:433-440).

    python examples/run_script_loop.py [script] [steps]          script: nmpc_tt (default, 700 steps like :341),
                                                                  t_trajectory, plus_trajectory, race_trajectory_1,
                                                                  race_track_2, 10_obstacles
CasADi is not needed (and not installed here); numpy stands in for ca.DM.
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import b200nmpc                      # noqa: E402


def shift_timestep(T, x0, u, xs, con_t):
    """NMPC_TT.py:13-30 with numpy: plant Euler step with the first input, warm start shifted by one stage (last
    repeated), target Euler step with (v, omega) = con_t."""
    v, th, ps = u[0, 0], x0[3], x0[4]
    f = np.array([v * np.cos(ps) * np.cos(th), v * np.sin(ps) * np.cos(th), v * np.sin(th), *u[1:, 0]])
    x0 = x0 + T * f
    u0 = np.concatenate([u[:, 1:], u[:, -1:]], axis=1)
    xs = xs + T * np.array([con_t[0] * np.cos(xs[2]), con_t[0] * np.sin(xs[2]), con_t[1]])
    return x0, u0, xs


def fov_centre(x0, vfov, hfov):
    """NMPC_TT.py:399-402."""
    a_p = (x0[2] * np.tan(x0[6] + vfov / 2) - x0[2] * np.tan(x0[6] - vfov / 2)) / 2
    b_p = (x0[2] * np.tan(x0[5] + hfov / 2) - x0[2] * np.tan(x0[5] - hfov / 2)) / 2
    return x0[0] + a_p + x0[2] * np.tan(x0[6] - vfov / 2), x0[1] + b_p + x0[2] * np.tan(x0[5] - hfov / 2)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "nmpc_tt"
    loop_run = int(sys.argv[2]) if len(sys.argv) > 2 else 700
    sc = b200nmpc.SCENARIOS[name]
    opts = {"ipopt": {"max_iter": 100, "print_level": 0, "acceptable_tol": 1e-8, "acceptable_obj_change_tol": 1e-6},
            "print_time": 0}                                                         # :257-265
    solver = b200nmpc.nlpsol("solver", "ipm", sc, opts)                              # <- the swapped line (:267)
    lbx, ubx, lbg, ubg = sc.bounds()                                                 # args['lbx'] ... (:269-313)
    x0 = np.array(sc.x_init, dtype=float); xs = np.array(sc.target_init, dtype=float)   # :320-326
    u0 = np.zeros((6, sc.N))                                                         # :329
    x_e = np.zeros(loop_run + 1); y_e = np.zeros(loop_run + 1); ss = np.zeros((3, loop_run + 1)); ss[:, 0] = xs
    conv = iters = 0
    t_start = time.perf_counter()
    for mpc_iter in range(loop_run):                                                 # while mpc_iter < loop_run (:348)
        p = np.concatenate([x0, xs])                                                 # args['p'] (:350-353)
        w0 = u0.T.reshape(-1)                                                        # reshape(u0, 6N, 1) (:356)
        sol = solver(x0=w0, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, p=p)                 # :358-365
        u = sol["x"].reshape(sc.N, 6).T                                              # ca.reshape(sol['x'], 6, N) (:367)
        st = solver.stats(); conv += int(st["success"][0]); iters += int(st["iter_count"][0])
        x0, u0, xs = shift_timestep(sc.T, x0, u, xs, sc.schedule(mpc_iter))  # :382
        x_e[mpc_iter + 1], y_e[mpc_iter + 1] = fov_centre(x0, sc.vfov, sc.hfov)   # :399-402
        ss[:, mpc_iter + 1] = xs
    total = time.perf_counter() - t_start
    error = np.hypot(x_e[1:loop_run + 1] - ss[0, :loop_run], y_e[1:loop_run + 1] - ss[1, :loop_run])   # :433-436
    print(f"{sc.script}: {loop_run} MPC steps in {total:.2f} s ({1e3 * total / loop_run:.2f} ms per step incl. host work), "
          f"{conv} converged, {iters / loop_run:.1f} IPM iterations per step")
    print("sum of FOV-centre errors:", float(error.sum()))                            # print(sum(error1[0:itr])) (:440)


if __name__ == "__main__":
    main()
