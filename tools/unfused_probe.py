"""why is solve + nmpc_step (two launches) slower than nmpc_solve_and_step under pipelining? (GPU)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import PipelinedClosedLoop
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = 4096; K, W, S = 8, 3, 8
for mode in ("fused", "unfused", "unfused+keep", "unfused+events", "fused+events"):
    p, vw = b200nmpc.random_instances(sc, B, seed=2000)
    pl = PipelinedClosedLoop(lambda n: b200nmpc.nlpsol('s', 'ipm', sc, max_batch=n), sc, p, target_vw=vw, pipelines=S)
    for k in range(W): pl.step()
    torch.cuda.synchronize()
    keep = []; evs = []
    t0 = time.perf_counter()
    for k in range(K):
        for lp, st in zip(pl.loops, pl.streams):
            with torch.cuda.stream(st):
                if "events" in mode:
                    e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
                if mode.startswith("fused"):
                    lp.solver.solve_and_step(lp.p, lp.u_warm, lp.lbx, lp.ubx, lp.lbg, lp.ubg, lp.vw, lp.fov, lp.err_sum)
                else:
                    sol = lp.solver(x0=lp.u_warm, p=lp.p, lbx=lp.lbx, ubx=lp.ubx, lbg=lp.lbg, ubg=lp.ubg, want_g=False, want_lam=False)
                    lp.solver.step(sol["x"], lp.p, lp.u_warm, lp.vw, lp.fov, lp.err_sum)
                if "keep" in mode:
                    keep.append(lp.solver._stats)
                if "events" in mode:
                    e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(e)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'{mode}: {1e3*(t2-t0)/K:.2f} ms/step (host enqueue {1e3*(t1-t0)/K:.2f} ms/step)')
