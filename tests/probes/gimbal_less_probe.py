"""model 1 (gimbal-less tracker) on the GPU against the oracle: function level, random solves, the script's closed loop"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import b200nmpc, oracle
from mpc_implementation_b200.closed_loop import ClosedLoop
sc = b200nmpc.SCENARIOS["gimbal_less"]
sp = oracle.make_spec(sc.T, sc.N, 0, model=1)
lbx, ubx, lbg, ubg = sc.bounds(); obs = np.zeros((0, 3))
B = 512
p, vw = b200nmpc.random_instances(sc, B, seed=3)
s = b200nmpc.nlpsol("s", "ipm", sc, max_batch=B)
rng = np.random.default_rng(0)
w = lbx + (ubx - lbx) * rng.uniform(size=(B, sc.n_w)); lam = rng.normal(size=(B, sc.n_g)); v = rng.normal(size=(B, sc.n_w))
e = {k: (v_.cpu().numpy() if torch.is_tensor(v_) else v_) for k, v_ in s.evaluate(w, p, lam=lam, v=v).items()}
for b in range(3):
    o = oracle.evaluate(sp, obs, w[b], p[b], lam_g=lam[b], hessian=True)
    print("eval", abs(o["f"] - float(e["f"][b])), np.abs(o["g"] - np.asarray(e["g"][b])).max(), np.abs(o["grad"] - np.asarray(e["grad"][b])).max(),
          np.abs(o["J"].T @ lam[b] - np.asarray(e["jtv"][b])).max(), np.abs(o["H"] @ v[b] - np.asarray(e["hv"][b])).max())
x0 = np.zeros((B, sc.n_w))
r = {k: (v_.cpu().numpy() if torch.is_tensor(v_) else v_) for k, v_ in s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg).items()}
st = s.stats()
o = oracle.solve(sp, obs, p, x0, lbx, ubx, lbg, ubg)
c = lambda a: a.cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
gs, gi = c(st["return_status"]), c(st["iter_count"])
print("status equal", (gs == o["status"]).mean(), "iters equal", (gi == o["iters"]).mean(), "converged", (gs == 0).mean(), (o["status"] == 0).mean())
ok = (gs == 0) & (o["status"] == 0)
print("max |dx|", np.abs(np.asarray(r["x"]) - o["x"])[ok].max(), "max rel df", (np.abs(np.asarray(r["f"]).ravel() - o["f"])[ok] / np.abs(o["f"][ok])).max(),
      "lam_g", np.abs(np.asarray(r["lam_g"]) - o["lam_g"])[ok].max())
# the script's own closed loop (NMPC_TT.m:138-170): 100 steps from x0, xs with u0 = 0
p0 = np.array([list(sc.x_init) + list(sc.target_init)])
cl = ClosedLoop(b200nmpc.nlpsol("c", "ipm", sc, max_batch=1), sc, p0)
from oracle import nlp_ref
x0_, xs_, u0_ = np.array(sc.x_init), np.array(sc.target_init), np.zeros((15, 3))
its = []
for k in range(sc.steps):
    cl.step()
    its.append(int(cl.solver.stats()["iter_count"][0]))
    ro = oracle.solve(sp, obs, np.concatenate([x0_, xs_])[None], u0_.reshape(1, -1), lbx, ubx, lbg, ubg, nthreads=1)
    x0_, u0_, xs_ = nlp_ref.shift1(sc.T, x0_, ro["x"].reshape(15, 3), xs_)
print("closed loop iters", its[:20], "final state GPU", cl.p.cpu().numpy(), "oracle", x0_, xs_)
