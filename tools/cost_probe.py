"""where does a closed-loop batch step go?  per-instance work of one step (iterations, line-search trials, iterations inside the
restoration phase, inertia retries) at step 20 of config 2, grouped by return status, and the wall time of that step when
only the instances of one group are solved (same p / warm starts)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import ClosedLoop
np.set_printoptions(linewidth=200, precision=3, suppress=True)
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = 4096
p, vw = b200nmpc.random_instances(sc, B, seed=2000)
s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
cl = ClosedLoop(s, sc, p, target_vw=vw)
for k in range(20): cl.step()
dbg = torch.zeros((B, 101, 10), dtype=torch.float64, device='cuda')
b200nmpc._ffi.lib().nmpc_set_debug_log(s._h, dbg.data_ptr(), 101)
pk, uk = cl.p.clone(), cl.u_warm.clone()
cl.step(); st = s.stats(); torch.cuda.synchronize()
b200nmpc._ffi.lib().nmpc_set_debug_log(s._h, None, 0)
sg = st['return_status'].cpu().numpy(); ig = st['iter_count'].cpu().numpy(); lg = dbg.cpu().numpy()
ls = lg[:, :, 7].sum(1); rs = lg[:, :, 9].sum(1); dwn = (lg[:, :, 4] > 0).sum(1)
print('status histogram', np.bincount(sg, minlength=7))
for c in range(7):
    m = sg == c
    if m.any(): print(f'status {c}: n={m.sum():5d} iters {ig[m].mean():6.1f} ls trials {ls[m].mean():7.1f} resto iters {rs[m].mean():6.1f} iterations with dw>0 {dwn[m].mean():6.1f}')
lbx, ubx, lbg, ubg = sc.bounds()
def time_subset(idx, reps=3):
    n = len(idx)
    if n == 0: return 0.0
    ss = b200nmpc.nlpsol('t', 'ipm', sc, max_batch=n)
    best = 1e9
    for r in range(reps):
        pp, uu = pk[idx].clone(), uk[idx].clone()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ss(x0=uu, p=pp, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, want_g=False, want_lam=False); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
allm = time_subset(torch.arange(B, device='cuda'))
print(f'whole batch {allm:.2f} ms')
for name, m in (('converged', sg == 0), ('not converged', sg != 0), ('iters >= 60', ig >= 60), ('iters < 60', ig < 60)):
    idx = torch.as_tensor(np.flatnonzero(m), device='cuda')
    print(f'{name:14s} n={int(m.sum()):5d}: {time_subset(idx):7.2f} ms alone;  iterations {int(ig[m].sum())}, ls trials {int(ls[m].sum())}')
# slowest single instances
order = np.argsort(-ig)[:6]
for i in order:
    t = time_subset(torch.as_tensor([i], device='cuda'))
    print(f'instance {i}: status {sg[i]} iters {ig[i]} ls {int(ls[i])} resto iters {int(rs[i])}: {t:.2f} ms alone ({1e3 * t / max(1, ig[i]):.0f} us per iteration)')
