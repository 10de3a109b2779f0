"""what are the status-2 (restoration needed) instances of the closed loop doing? (GPU)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import ClosedLoop
np.set_printoptions(linewidth=200, precision=4, suppress=True)
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = 4096
p, vw = b200nmpc.random_instances(sc, B, seed=2000)
s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
cl = ClosedLoop(s, sc, p, target_vw=vw)
hist = []
for k in range(22):
    cl.step(); st = s.stats(); wc = s.work_counters()
    sg = st['return_status'].cpu().numpy(); ig = st['iter_count'].cpu().numpy()
    hist.append(sg.copy())
    print(k, 'status', np.bincount(sg, minlength=4)[:4], 'iters by status', [round(float(ig[sg == c].mean()), 1) if (sg == c).any() else 0 for c in range(3)], 'ls/iter', round(wc['ls_trials'] / ig.sum(), 2), 'fact/iter', round(wc['factorizations'] / ig.sum(), 2))
hist = np.array(hist)
stuck = np.where(hist[-1] == 2)[0]
print('stuck now', len(stuck), ' of which stuck also 5 steps ago', int((hist[-6][stuck] == 2).sum()), ' ever converged after first failure:', int(sum(((hist[:, i] != 0).argmax() < 21) and (hist[(hist[:, i] != 0).argmax():, i] == 0).any() for i in stuck)))
# log one more step
dbg = torch.zeros((B, 101, 8), dtype=torch.float64, device='cuda')
b200nmpc._ffi.lib().nmpc_set_debug_log(s._h, dbg.data_ptr(), 101)
pk = cl.p.clone(); cl.step(); st = s.stats()
sg = st['return_status'].cpu().numpy(); ig = st['iter_count'].cpu().numpy()
lg = dbg.cpu().numpy()
for i in list(np.where(sg == 1)[0][:5]):
    print('--- instance', i, 'iters', ig[i], 'p', pk[i].cpu().numpy())
    print('   g0 rows: z', float(pk[i, 2]), 'theta', float(pk[i, 3]), 'X5..7', pk[i, 5:8].cpu().numpy())
    n = ig[i]
    print('   [mu, f, inf_pr, inf_du, dw, a_pr, a_du, ls]')
    print(lg[i, 0:n:6]); print(lg[i, n-4:n])
ls_by = [lg[sg == c, :, 7].sum() / max(1, (sg == c).sum()) for c in range(3)]
print('ls trials per instance by status', ls_by)
