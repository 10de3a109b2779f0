"""status / iteration histogram of the bench workload's warm closed-loop steps (GPU) and of the oracle on the same instances"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import b200nmpc, oracle
from mpc_implementation_b200.closed_loop import ClosedLoop
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = 2048
p, vw = b200nmpc.random_instances(sc, B, seed=2000)
s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
cl = ClosedLoop(s, sc, p, target_vw=vw)
osp = oracle.make_spec(sc.T, sc.N, sc.n_obs); obs = sc.obstacle_table(); lbx, ubx, lbg, ubg = sc.bounds()
for k in range(6):
    pk = cl.p.cpu().numpy().copy(); uk = cl.u_warm.cpu().numpy().copy()
    cl.step(); st = s.stats()
    sg = st['return_status'].cpu().numpy(); ig = st['iter_count'].cpu().numpy()
    if k >= 3:
        ro = oracle.solve(osp, obs, pk[:512], uk[:512], lbx, ubx, lbg, ubg, want_g=False, want_lam=False)
        print('step', k, 'GPU status', np.bincount(sg, minlength=6), 'iters by status', [int(ig[sg == c].mean()) if (sg == c).any() else 0 for c in range(6)],
              '| oracle(512) status', np.bincount(ro['status'], minlength=6), 'iters by status', [int(ro['iters'][ro['status'] == c].mean()) if (ro['status'] == c).any() else 0 for c in range(6)])
        both = np.stack([ro['status'], sg[:512]], 1)
        print('   confusion (oracle,gpu):', {(a, b): int(((both[:, 0] == a) & (both[:, 1] == b)).sum()) for a in range(4) for b in range(4) if ((both[:, 0] == a) & (both[:, 1] == b)).any()})
