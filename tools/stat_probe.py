"""why do a few converged instances have a large unscaled stationarity residual? (GPU)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc as pkg
sc = pkg.SCENARIOS['t_trajectory']; B = 65536; dev = 'cuda:0'
lbx, ubx, lbg, ubg = sc.bounds()
p, _ = pkg.random_instances(sc, B, seed=3015)
T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
x0 = T(np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N))); pt = T(p)
s = pkg.nlpsol('s', 'ipm', sc, max_batch=B)
sol = s(x0=x0, p=pt, lbx=T(lbx), ubx=T(ubx), lbg=T(lbg), ubg=T(ubg))
st = s.stats(); ok = st['success']
ev = s.evaluate(sol['x'], pt, lam=sol['lam_g'])
ev0 = s.evaluate(x0, pt)
df = torch.clamp(100.0 / ev0['grad'].abs().amax(dim=1), max=1.0)
r = (ev['grad'] + ev['jtv'] + sol['lam_x']).abs().amax(dim=1)
mult = torch.maximum(sol['lam_g'].abs().amax(dim=1), sol['lam_x'].abs().amax(dim=1)).clamp(min=1.0)
gscale = torch.maximum(ev['grad'].abs().amax(dim=1), mult)
ratio = r / (1e-6 + 2e-6 * gscale); ratio[~ok] = 0
idx = torch.argsort(ratio, descending=True)[:8]
n_mult = sc.n_g + sc.n_w
for i in idx.tolist():
    sd = max(100.0, float(df[i]) * (sol['lam_g'][i].abs().sum() + sol['lam_x'][i].abs().sum()) / n_mult) / 100.0
    print(f"inst {i}: ratio {ratio[i]:.1f} r {r[i]:.3e} gscale {gscale[i]:.3e} df {df[i]:.3e} r*df {r[i]*df[i]:.3e} sd~{sd:.2f} iters {int(st['iter_count'][i])} max|lam_g| {sol['lam_g'][i].abs().max():.3e} max|lam_x| {sol['lam_x'][i].abs().max():.3e} f {sol['f'][i]:.4f}")
    # which component?
    rv = (ev['grad'] + ev['jtv'] + sol['lam_x'])[i]; j = int(rv.abs().argmax())
    xi = sol['x'][i, j]; print(f"     worst comp w[{j}] (stage {j//6}, ctl {j%6}) = {xi:.12f} bounds [{lbx[j]}, {ubx[j]}] lam_x {sol['lam_x'][i, j]:.3e} grad {ev['grad'][i, j]:.3e} jtv {ev['jtv'][i, j]:.3e}")
