"""small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): a few instances of two
instantiations through solve, eval and step"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc
from mpc_implementation_b200.closed_loop import ClosedLoop
for name, B in (("nmpc_tt", 20), ("race_track_2", 6)):
    sc = b200nmpc.SCENARIOS[name]
    p, vw = b200nmpc.random_instances(sc, B, seed=5)
    s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
    cl = ClosedLoop(s, sc, p, target_vw=vw, predict_target=(name == "nmpc_tt"))
    for k in range(2):
        sol = cl.step(want_g=True, want_lam=True)
    torch.cuda.synchronize()
    ev = s.evaluate(sol["x"], cl.p, lam=sol["lam_g"], v=sol["x"])
    print(name, s.stats()["return_status"].cpu().numpy(), float(ev["f"].sum()))
print("done")
