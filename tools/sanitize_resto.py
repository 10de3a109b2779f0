"""Small end-to-end case that drives the rare paths (restoration phase, watchdog, soft restoration, slack repair,
device-side schedule) of two instantiations -- for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tools/sanitize_resto.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import ClosedLoop

for name, N in (("nmpc_tt", 15), ("race_track_2", 30), ("10_obstacles", 15)):
    sc = b200nmpc.SCENARIOS[name]
    if N != sc.N:
        sc = sc.with_horizon(N)
    B = 24
    p, vw = b200nmpc.random_instances(sc, B, seed=99)
    ob = sc.obstacle_table()
    p[:8, 0] = ob[1, 0] + 10.0; p[:8, 1] = ob[1, 1] - 5.0          # inside an obstacle: infeasible, restoration phase
    p[8:12, 2] = 150.0 + 1e-4                                      # above the ceiling
    s = b200nmpc.nlpsol("s", "ipm", sc, max_batch=B)
    cl = ClosedLoop(s, sc, p, target_vw=None if name != "nmpc_tt" else vw, phase=None if name == "nmpc_tt" else np.arange(B) * 7)
    for k in range(4):
        cl.step()
    torch.cuda.synchronize()
    st = s.stats()
    print(name, N, "status", np.bincount(st["return_status"].cpu().numpy(), minlength=7).tolist(), s.work_counters())
