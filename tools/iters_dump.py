"""dump per-instance iteration counts / status of consecutive warm closed-loop steps of the bench workload (GPU)
to gpurun_out/iters.npz, for offline study of the scheduling tail"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import ClosedLoop
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = 4096
p, vw = b200nmpc.random_instances(sc, B, seed=1234)
s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
cl = ClosedLoop(s, sc, p, target_vw=vw)
its, sts, tms = [], [], []
for k in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cl.step(); e1.record(); torch.cuda.synchronize()
    st = s.stats(); wc = s.work_counters()
    its.append(st['iter_count'].cpu().numpy().copy()); sts.append(st['return_status'].cpu().numpy().copy()); tms.append(e0.elapsed_time(e1))
    print(k, tms[-1], np.bincount(sts[-1], minlength=6), its[-1].mean(), wc)
Path('gpurun_out').mkdir(exist_ok=True)
np.savez('gpurun_out/iters.npz', iters=np.array(its), status=np.array(sts), ms=np.array(tms))
