"""GPU vs oracle on the hard (restoration / watchdog) instances of a closed loop: status confusion, and the first
iteration at which the per-iteration logs of the two implementations part (kernel debug log vs oracle.solve_log).
Usage: python tests/probes/resto_debug.py [scenario] [B] [steps] [max_show]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import b200nmpc, oracle

scn = sys.argv[1] if len(sys.argv) > 1 else "nmpc_tt"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
show = int(sys.argv[4]) if len(sys.argv) > 4 else 6
sc = b200nmpc.SCENARIOS[scn]
lbx, ubx, lbg, ubg = sc.bounds()
sp = oracle.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov); obs = sc.obstacle_table()
p, vw = b200nmpc.random_instances(sc, B, seed=2000)
u = np.zeros((B, sc.n_w))
s = b200nmpc.nlpsol("s", "ipm", sc, max_batch=B)
ROWS = 101
dbg = torch.zeros((B, ROWS, 10), dtype=torch.float64, device="cuda")
L = b200nmpc._ffi.lib()


def shift(p, x):
    T = sc.T; u0 = x[:, :6]; th, ps = p[:, 3], p[:, 4]; pn = p.copy()
    pn[:, 0] += T * u0[:, 0] * np.cos(ps) * np.cos(th); pn[:, 1] += T * u0[:, 0] * np.sin(ps) * np.cos(th); pn[:, 2] += T * u0[:, 0] * np.sin(th)
    pn[:, 3:8] += T * u0[:, 1:6]
    tt = p[:, 10]
    pn[:, 8] += T * vw[:, 0] * np.cos(tt); pn[:, 9] += T * vw[:, 0] * np.sin(tt); pn[:, 10] += T * vw[:, 1]
    return pn, np.concatenate([x[:, 6:], x[:, -6:]], axis=1)


for k in range(steps):
    dbg.zero_()
    L.nmpc_set_debug_log(s._h, dbg.data_ptr(), ROWS)
    sol = s(x0=u, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    L.nmpc_set_debug_log(s._h, None, 0)
    st = s.stats(); sg, ig = st["return_status"], st["iter_count"]
    ref = oracle.solve(sp, obs, p, u, lbx, ubx, lbg, ubg, want_g=False, want_lam=False)
    so, io = ref["status"], ref["iters"]
    conf = {}
    for a, b in zip(so, sg):
        conf[(int(a), int(b))] = conf.get((int(a), int(b)), 0) + 1
    both = (so == 0) & (sg == 0)
    df = np.abs(ref["f"][both] - sol["f"][both]) / np.maximum(1, np.abs(ref["f"][both]))
    print(f"step {k}: confusion (oracle, gpu) {conf}  iters equal {(io == ig).mean():.3f}  max rel f diff (both ok) {df.max() if both.any() else 0:.2e}"
          f"  gpu counters {s.work_counters()}", flush=True)
    hard = np.where((so != sg) | ((io != ig) & ((ref['stats'][:, 2:7].sum(1) > 0))))[0]
    D = dbg.cpu().numpy()
    for b in hard[:show]:
        r = oracle.solve_log(sp, obs, p[b], u[b], lbx, ubx, lbg, ubg, max_log=256)
        lo = r["log"]; lg = D[b]
        print(f"  inst {b}: oracle status {so[b]} it {io[b]} stats {ref['stats'][b].tolist()} | gpu status {sg[b]} it {ig[b]}")
        # oracle logs restoration iterations BEFORE the outer iteration that called them; re-order by iteration index:
        # outer line of the calling iteration first (tag R), then the inner ones
        seq = []
        i = 0
        while i < len(lo):
            if lo[i][8] >= 1000:
                j = i
                while j < len(lo) and lo[j][8] >= 1000:
                    j += 1
                if j < len(lo):
                    seq.append(lo[j])
                seq.extend(lo[i:j]); i = j + 1
            else:
                seq.append(lo[i]); i += 1
        n = min(len(seq), int(ig[b]) + 1, ROWS)
        first = None
        for it in range(n):
            a, g = seq[it], lg[it]
            if g[0] == 0:
                break
            rel = lambda x, y: abs(x - y) / max(1e-300, abs(x), abs(y))
            if rel(a[0], g[0]) > 1e-6 or rel(a[5], g[5]) > 1e-4 or int(a[7]) != int(g[7]) or (int(a[8]) % 1000) != int(g[8]):
                first = it; break
        print(f"    first divergence at iteration {first}")
        fmt = lambda a, mode, tag: "mu %.3e f %.6e pr %.3e du %.3e dw %.2e a %.4e/%.4e ls %2d %s%s" % (a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], 'R' if mode else ' ', chr(int(tag)) if tag > 0 else '?')
        print("    oracle tail:")
        for a in seq[-5:]:
            print("         " + fmt(a, a[8] >= 1000, int(a[8]) % 1000))
        print("    gpu tail:")
        for it in range(max(0, int(ig[b]) - 5), min(int(ig[b]) + 1, ROWS)):
            g = lg[it]
            if g[0] != 0:
                print("     %3d " % it + fmt(g, g[9], g[8]))
        lo_i = max(0, (first or 0) - 3)
        for it in range(lo_i, min(n, lo_i + 8)):
            a, g = seq[it], lg[it]
            print("     %3d O mu %.3e f %.6e pr %.3e du %.3e dw %.2e a %.4e/%.4e ls %2d %s%s" % (it, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], 'R' if a[8] >= 1000 else ' ', chr(int(a[8]) % 1000)))
            print("         G mu %.3e f %.6e pr %.3e du %.3e dw %.2e a %.4e/%.4e ls %2d %s%s" % (g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7], 'R' if g[9] else ' ', chr(int(g[8])) if g[8] > 0 else '?'))
    p, u = shift(p, ref["x"])      # teacher-forced by the oracle
