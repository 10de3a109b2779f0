import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import b200nmpc, oracle
from mpc_implementation_b200.closed_loop import ClosedLoop
np.set_printoptions(linewidth=250, precision=5)
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = 512
p, vw = b200nmpc.random_instances(sc, B, seed=2000)
s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
cl = ClosedLoop(s, sc, p, target_vw=vw)
osp = oracle.make_spec(sc.T, sc.N, sc.n_obs); obs = sc.obstacle_table(); lbx, ubx, lbg, ubg = sc.bounds()
for k in range(4):
    pk = cl.p.cpu().numpy().copy(); uk = cl.u_warm.cpu().numpy().copy()
    cl.step()
sg = s.stats()['return_status'].cpu().numpy()
ro = oracle.solve(osp, obs, pk, uk, lbx, ubx, lbg, ubg, want_g=False, want_lam=False)
for code in (2, 3):
    idx = np.where((ro['status'] == 0) & (sg == code))[0][:2]
    for b in idx:
        dbg = torch.zeros((1, 101, 8), dtype=torch.float64, device='cuda')
        b200nmpc._ffi.lib().nmpc_set_debug_log(s._h, dbg.data_ptr(), 101)
        r1 = s(x0=uk[b], p=pk[b], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg); st = s.stats()
        b200nmpc._ffi.lib().nmpc_set_debug_log(s._h, None, 0)
        lo = oracle.solve_log(osp, obs, pk[b], uk[b], lbx, ubx, lbg, ubg)['log']; lg = dbg.cpu().numpy()[0]
        n = int(st['iter_count'][0])
        print(f'== instance {b}: gpu status {int(st["return_status"][0])} iters {n}; oracle iters {len(lo)}  [mu, f, inf_pr, inf_du, dw, alpha_pr, alpha_du, ls]')
        for i in range(max(0, min(n, len(lo)) - 7), max(n, len(lo))):
            if i < len(lo): print(i, 'O', lo[i])
            if i < n: print(i, 'G', lg[i])
