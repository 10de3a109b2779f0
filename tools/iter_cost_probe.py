"""Kernel time per IPM iteration-instance on a well-behaved batch (one solve of B instances, no closed loop):
   NMPC_B200_LIB=<lib> python tools/iter_cost_probe.py [scenario] [B] [N]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
scn = sys.argv[1] if len(sys.argv) > 1 else "t_trajectory"; B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
sc = b200nmpc.SCENARIOS[scn]
if len(sys.argv) > 3 and int(sys.argv[3]) != sc.N:
    sc = sc.with_horizon(int(sys.argv[3]))
lbx, ubx, lbg, ubg = sc.bounds()
p, _ = b200nmpc.random_instances(sc, B, seed=7)
x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))
dev = "cuda"
t = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
s = b200nmpc.nlpsol("s", "ipm", sc, max_batch=B)
P, X0 = t(p), t(x0)
for rep in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); sol = s(x0=X0, p=P, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, want_g=False, want_lam=False); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); st = s.stats(); it = st["iter_count"].sum().item(); ok = (st["return_status"] == 0).float().mean().item()
    wc = s.work_counters()
    print(f"{scn} N={sc.N} B={B}: {ms:.2f} ms, {it} iterations, {ms * 1e6 / it:.1f} ns per iteration-instance, converged {ok:.4f}, max it {st['iter_count'].max().item()}, "
          f"fact/iter {wc['factorizations'] / it:.2f} trials/iter {wc['ls_trials'] / it:.2f}", flush=True)
