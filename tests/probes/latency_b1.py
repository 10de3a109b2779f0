"""single-instance latency (BASELINE config 1: batch 1) through the host-buffer call and on the device"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import b200nmpc, oracle
for name in ("t_trajectory", "nmpc_tt"):
    sc = b200nmpc.SCENARIOS[name]
    lbx, ubx, lbg, ubg = sc.bounds()
    s = b200nmpc.nlpsol("solver", "ipm", sc)
    p0 = np.array(list(sc.x_init) + list(sc.target_init))
    # a warm problem: a few closed-loop steps first
    from mpc_implementation_b200.closed_loop import ClosedLoop
    cl = ClosedLoop(s, sc, p0)
    for _ in range(12): cl.step()
    torch.cuda.synchronize()
    p = cl.p.cpu().numpy()[0].copy(); u = cl.u_warm.cpu().numpy()[0].copy()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); sol = s(x0=u, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg); ts.append(time.perf_counter() - t0)
    it = int(s.stats()["iter_count"][0]); st = int(s.stats()["return_status"][0])
    sp = oracle.make_spec(sc.T, sc.N, sc.n_obs)
    t0 = time.perf_counter(); r = oracle.solve(sp, sc.obstacle_table(), p, u, lbx, ubx, lbg, ubg, nthreads=1); tcpu = time.perf_counter() - t0
    print(f"{name}: warm single solve, {it} iterations (status {st}): host call p50 {1e3*np.median(ts):.2f} ms (min {1e3*min(ts):.2f}); "
          f"CPU oracle, one thread: {1e3*tcpu:.1f} ms ({int(r['iters'][0])} iterations)")
