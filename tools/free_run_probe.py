"""free-running closed loop (nmpc_run_closed_loop) against the per-step loops: converged solves/s
   python tools/free_run_probe.py [config] [steps]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc, bench
from mpc_implementation_b200.closed_loop import ClosedLoop, PipelinedClosedLoop
cfgi = int(sys.argv[1]) if len(sys.argv) > 1 else 2; K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = bench.CONFIGS[cfgi]; B = cfg["batch"]
sc, p, vw, obs, sched = bench.make_workload(b200nmpc, cfg, B, seed=1000)
kw = dict(target_vw=vw) if sched is None else dict(schedules=sched["schedules"], schedule_of=sched["schedule_of"], phase=sched["phase"])
ob = None if obs is None else torch.as_tensor(obs, device="cuda")
def timed(fn):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1), r
W = 5
# (a) one batch, one launch per step
a = ClosedLoop(b200nmpc.nlpsol("a", "ipm", sc, max_batch=B), sc, p, obstacles=ob, **kw)
for _ in range(W): a.step()
def run_a():
    c = torch.zeros((), dtype=torch.int64, device="cuda")
    for _ in range(K): a.step(); c += a.solver.stats()["success"].sum()
    return int(c.item())
ms, conv = timed(run_a); print(f"config {cfgi} B={B} K={K}: per-step single batch   {ms / K:8.2f} ms/step {conv / ms * 1e3:10.0f} solves/s  conv {conv / (B * K):.4f}", flush=True)
# (b) free running, chunks of K steps
b = ClosedLoop(b200nmpc.nlpsol("b", "ipm", sc, max_batch=B), sc, p, obstacles=ob, **kw)
b.run_free(W, log=False)
for chunk in (K, max(1, K // 4)):
    def run_b():
        c = torch.zeros((), dtype=torch.int64, device="cuda")
        for _ in range(K // chunk): c += b.run_free(chunk, log=False)["converged"].sum()
        return int(c.item())
    ms, conv = timed(run_b); n = (K // chunk) * chunk
    print(f"config {cfgi} B={B} K={n}: free running, {chunk:3d} steps/launch {ms / n:8.2f} ms/step {conv / ms * 1e3:10.0f} solves/s  conv {conv / (B * n):.4f}", flush=True)
