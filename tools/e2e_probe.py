"""where does the host-buffer path spend its time? (GPU)"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from bench import host_shift
from mpc_implementation_b200.closed_loop import ClosedLoop
sc = b200nmpc.SCENARIOS['nmpc_tt']; B = 4096
p, vw = b200nmpc.random_instances(sc, B, seed=2000)
s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
cl = ClosedLoop(s, sc, p, target_vw=vw)
for k in range(8): cl.step()
torch.cuda.synchronize()
ph = torch.empty((B, 11), dtype=torch.float64).pin_memory().numpy(); ph[:] = cl.p.cpu().numpy()
uh = torch.empty((B, sc.n_w), dtype=torch.float64).pin_memory().numpy(); uh[:] = cl.u_warm.cpu().numpy()
lbx, ubx, lbg, ubg = sc.bounds()
for k in range(6):
    # device path on the same inputs (kernel time by events)
    pd, ud = torch.from_numpy(ph).cuda(), torch.from_numpy(uh).cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sd = s(x0=ud, p=pd, lbx=cl.lbx, ubx=cl.ubx, lbg=cl.lbg, ubg=cl.ubg, want_g=False, want_lam=False); e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    s2 = s(x0=uh, p=ph, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, want_g=False, want_lam=False)
    t1 = time.perf_counter()
    it = s.stats()['iter_count']
    uh[:] = host_shift(sc.T, ph, s2["x"], vw)
    t2 = time.perf_counter()
    print(f'step {k}: device-path kernel {e0.elapsed_time(e1):.2f} ms | host call {1e3*(t1-t0):.2f} ms | host_shift {1e3*(t2-t1):.2f} ms | iters mean {it.mean():.2f} max {it.max()} same x: {np.abs(sd["x"].cpu().numpy()-s2["x"]).max():.2e}')
