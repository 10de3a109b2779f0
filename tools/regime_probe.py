"""iterations per second of the bulk regime (B = 16384, one batch) for two scenarios with different inertia-retry rates:
how much of the lockstep waiting comes from repeated factorisations?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import ClosedLoop
B = 16384
for name in ("nmpc_tt", "t_trajectory", "race_track_2"):
    sc = b200nmpc.SCENARIOS[name]
    p, vw = b200nmpc.random_instances(sc, B, seed=7)
    s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
    cl = ClosedLoop(s, sc, p, target_vw=vw)
    for _ in range(4): cl.step()
    torch.cuda.synchronize()
    tot_it = tot_f = tot_ls = 0; ms = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cl.step(); e1.record(); torch.cuda.synchronize()
        wc = s.work_counters(); st = s.stats()
        tot_it += int(st["iter_count"].sum()); tot_f += wc["factorizations"]; tot_ls += wc["ls_trials"]; ms += e0.elapsed_time(e1)
    print(f'{name}: {ms/4:.1f} ms/step, mean iters {tot_it/4/B:.1f}, fact/iter {tot_f/tot_it:.2f}, ls/iter {tot_ls/tot_it:.2f}, '
          f'{tot_it/ms/1e3:.2f} M iterations/s, {tot_f/ms/1e3:.2f} M factorisations/s, conv {float(st["success"].double().mean()):.3f}')
