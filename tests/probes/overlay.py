"""Free-running closed loops of the reference scripts: the GPU loop (nmpc_solve + nmpc_step, one instance) next to the
CPU oracle loop (oracle solve + the restated shift_timestep), started from the scripts' own initial conditions and
target schedules.  Writes profiles/overlay_<script>.csv (step, UAV x y z, FOV centre, target, per side) and prints
the largest deviation -- the "closed-loop trajectories overlaid" of the north star, as numbers (no matplotlib here)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import b200nmpc, oracle
from oracle import nlp_ref
from mpc_implementation_b200.closed_loop import ClosedLoop

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 120
for name in ("nmpc_tt", "t_trajectory", "plus_trajectory", "race_track_2"):
    sc = b200nmpc.SCENARIOS[name]
    lbx, ubx, lbg, ubg = sc.bounds()
    s = b200nmpc.nlpsol("solver", "ipm", sc)
    p0 = np.array(list(sc.x_init) + list(sc.target_init))
    cl = ClosedLoop(s, sc, p0)
    sp = oracle.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    po = p0.copy(); wo = np.zeros(sc.n_w)
    rows = []; dev_state = 0.0; dev_u0 = 0.0; errs = [0.0, 0.0]; conv = [0, 0]; its = [0, 0]; first = None; unconv_before = 0
    for i in range(STEPS):
        tgt = cl.p[0, 8:10].cpu().numpy().copy()
        sol = cl.step(); st = s.stats()
        conv[0] += int(st["return_status"][0] == 0); its[0] += int(st["iter_count"][0])
        r = oracle.solve(sp, sc.obstacle_table(), po, wo, lbx, ubx, lbg, ubg, nthreads=1)
        conv[1] += int(r["status"][0] == 0); its[1] += int(r["iters"][0])
        x0n, u0n, xsn = nlp_ref.shift_timestep(sc.T, po[:8], r["x"][0].reshape(sc.N, 6).T, po[8:], sc.schedule(i))
        xe, ye = nlp_ref.fov_centre(x0n)
        errs[1] += float(np.hypot(xe - po[8], ye - po[9]))
        dev_u0 = max(dev_u0, float(np.abs(r["x"][0][:6] - sol["x"][0, :6].cpu().numpy()).max()))
        po = np.concatenate([x0n, xsn]); wo = u0n.T.reshape(-1)
        g = cl.p[0].cpu().numpy(); f = cl.fov[0].cpu().numpy()
        dstep = float(np.abs(g[:8] - po[:8]).max())
        if first is None and dstep > 1e-5:
            first = (i, int(st["return_status"][0]), int(r["status"][0]), unconv_before)
        unconv_before += int(st["return_status"][0] != 0 or r["status"][0] != 0)
        dev_state = max(dev_state, dstep)
        rows.append([i, *g[:3], *f, *tgt, *po[:3], xe, ye])
    errs[0] = float(cl.err_sum[0])
    out = ROOT / "profiles" / f"overlay_{name}.csv"
    np.savetxt(out, np.array(rows), delimiter=",", fmt="%.9g",
               header="step,gpu_x,gpu_y,gpu_z,gpu_fov_x,gpu_fov_y,target_x,target_y,oracle_x,oracle_y,oracle_z,oracle_fov_x,oracle_fov_y", comments="")
    print(f"{name}: {STEPS} steps | converged GPU {conv[0]} oracle {conv[1]} | mean iters GPU {its[0]/STEPS:.1f} oracle {its[1]/STEPS:.1f} | "
          f"max |state_gpu - state_oracle| {dev_state:.3e} | max |u0 diff| {dev_u0:.3e} | sum FOV error GPU {errs[0]:.6f} oracle {errs[1]:.6f} | "
          + ("loops agree to 1e-5 throughout" if first is None else
             f"first deviation > 1e-5 at step {first[0]} (status GPU {first[1]}, oracle {first[2]}; {first[3]} non-converged solves before it)"))
