"""soak run: many pipelined closed-loop steps on every kernel instantiation (looks for hangs / races in the alignment
barrier, the in-kernel bookkeeping for the next call and the fetch order); prints a checksum per configuration"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import ClosedLoop, PipelinedClosedLoop
cases = [("nmpc_tt", None, 4096, 8, 300), ("t_trajectory", None, 3000, 5, 150), ("race_track_2", None, 1500, 3, 100),
         ("race_track_2", 30, 700, 2, 60), ("t_trajectory", 5, 2000, 4, 150), ("10_obstacles", None, 999, 7, 100)]
for name, N, B, S, K in cases:
    sc = b200nmpc.SCENARIOS[name]
    if N: sc = sc.with_horizon(N)
    p, vw = b200nmpc.random_instances(sc, B, seed=B + K)
    pl = PipelinedClosedLoop(lambda n: b200nmpc.nlpsol('s', 'ipm', sc, max_batch=n, fill=2), sc, p, target_vw=vw, pipelines=S)
    one = ClosedLoop(b200nmpc.nlpsol('s1', 'ipm', sc, max_batch=B), sc, p, target_vw=vw)
    t0 = time.perf_counter()
    for k in range(K):
        pl.step()
    pl.synchronize(); t1 = time.perf_counter()
    for k in range(K):
        one.step()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    same = torch.equal(pl.p, one.p) and torch.equal(pl.err_sum, one.err_sum)
    st = pl.stats()
    print(f'{sc.script} N={sc.N} B={B} S={S} K={K}: pipelined {1e3*(t1-t0)/K:.2f} ms/step, single {1e3*(t2-t1)/K:.2f} ms/step, '
          f'identical={same}, converged last step {float(st["success"].double().mean()):.3f}, err checksum {float(pl.err_sum.sum()):.6f}')
    assert same
print("soak ok")
