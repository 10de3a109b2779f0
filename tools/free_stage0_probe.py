"""does opening the bounds of the constant stage-0 rows (they depend on p only) cure the degenerate instances? (GPU)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import ClosedLoop
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = 4096
for free0 in (False, True):
    p, vw = b200nmpc.random_instances(sc, B, seed=2000)
    s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
    cl = ClosedLoop(s, sc, p, target_vw=vw)
    if free0:
        R = 5 + sc.n_obs
        cl.lbg[:R] = -float('inf'); cl.ubg[:R] = float('inf')
    tot = 0.0
    for k in range(23):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cl.step(); e1.record(); torch.cuda.synchronize()
        st = s.stats(); sg = st['return_status'].cpu().numpy(); ig = st['iter_count'].cpu().numpy()
        if k >= 3: tot += e0.elapsed_time(e1)
        if k in (0, 1, 2, 5, 10, 15, 22):
            print('free0' if free0 else 'ref  ', k, 'ms', round(e0.elapsed_time(e1), 1), 'status', np.bincount(sg, minlength=4)[:4], 'mean iters', round(float(ig.mean()), 1))
    print('   mean ms over steps 3..22:', round(tot / 20, 2), ' tracking error sum (mean over instances):', round(float(cl.err_sum.mean()), 2))
