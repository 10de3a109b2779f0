"""GPU-side debugging aid (not a test): function-level and solve-level comparison with the CPU oracle."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import b200nmpc, oracle
np.set_printoptions(linewidth=220, precision=5)

def fn_level(name, B=64, seed=0):
    sc = b200nmpc.SCENARIOS[name]
    rng = np.random.default_rng(seed)
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = b200nmpc.random_instances(sc, B, seed)
    w = lbx + rng.random((B, sc.n_w)) * (ubx - lbx)
    lam = rng.standard_normal((B, sc.n_g)); v = rng.standard_normal((B, sc.n_w))
    s = b200nmpc.nlpsol('s', 'ipm', sc)
    r = s.evaluate(w, p, lam=lam, v=v, sigma=0.7)
    osp = oracle.make_spec(sc.T, sc.N, sc.n_obs); obs = sc.obstacle_table()
    err = dict(f=0, g=0, grad=0, jtv=0, hv=0)
    for b in range(B):
        o = oracle.evaluate(osp, obs, w[b], p[b], lam[b], 0.7, hessian=True)
        rel = lambda a, c: float(np.abs(a - c).max() / max(1.0, np.abs(c).max()))
        err['f'] = max(err['f'], rel(r['f'][b].cpu().numpy(), o['f']))
        err['g'] = max(err['g'], rel(r['g'][b].cpu().numpy(), o['g']))
        err['grad'] = max(err['grad'], rel(r['grad'][b].cpu().numpy(), o['grad']))
        err['jtv'] = max(err['jtv'], rel(r['jtv'][b].cpu().numpy(), o['J'].T @ lam[b]))
        err['hv'] = max(err['hv'], rel(r['hv'][b].cpu().numpy(), o['H'] @ v[b]))
    print(name, 'function-level max rel err', err)

def solve_level(name, B=32, seed=1, warm=False, show=1):
    sc = b200nmpc.SCENARIOS[name]
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = b200nmpc.random_instances(sc, B, seed)
    p[0] = np.array(list(sc.x_init) + list(sc.target_init))
    x0 = np.zeros((B, sc.n_w))
    osp = oracle.make_spec(sc.T, sc.N, sc.n_obs); obs = sc.obstacle_table()
    t = time.time(); ro = oracle.solve(osp, obs, p, x0, lbx, ubx, lbg, ubg); tc = time.time() - t
    s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
    L = b200nmpc._ffi.lib()
    dbg = torch.zeros((B, 101, 8), dtype=torch.float64, device='cuda')
    L.nmpc_set_debug_log(s._h, dbg.data_ptr(), 101)
    t = time.time(); rg = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg); tg = time.time() - t
    st = s.stats()
    print(f'== {name} B={B}: oracle {tc:.3f}s gpu {tg:.3f}s  counters', s.work_counters())
    print('oracle status', np.bincount(ro['status'], minlength=6), 'iters mean', ro['iters'].mean())
    print('gpu    status', np.bincount(st['return_status'], minlength=6), 'iters mean', st['iter_count'].mean())
    same = (ro['status'] == st['return_status'])
    both = (ro['status'] == 0) & (st['return_status'] == 0)
    fo, fg = ro['f'], rg['f']
    relf = np.abs(fo - fg) / np.maximum(1, np.abs(fo))
    u0o, u0g = ro['x'][:, :6], rg['x'][:, :6]
    relu = np.abs(u0o - u0g).max(axis=1) / np.maximum(1e-12, np.abs(u0o).max(axis=1))
    print('status equal', same.sum(), '/', B, ' both converged', both.sum(), ' iters equal', (ro['iters'] == st['iter_count']).sum())
    if both.any():
        print('  f rel err max (both conv)', relf[both].max(), ' u0 rel err max', relu[both].max(), ' lam_g abs err max',
              np.abs(ro['lam_g'] - rg['lam_g'])[both].max(), 'x err', np.abs(ro['x'] - rg['x'])[both].max())
    bad = np.where(~both | (relf > 1e-8) | (ro['iters'] != st['iter_count']))[0][:show]
    lg = dbg.cpu().numpy()
    for b in bad:
        print(f'-- instance {b}: oracle st {ro["status"][b]} it {ro["iters"][b]} f {fo[b]:.10f} | gpu st {st["return_status"][b]} it {st["iter_count"][b]} f {fg[b]:.10f}')
        lo = oracle.solve_log(osp, obs, p[b], x0[b], lbx, ubx, lbg, ubg)['log']
        n = max(len(lo), int(st['iter_count'][b]))
        for i in range(min(n, 40)):
            a = lo[i] if i < len(lo) else np.zeros(8)
            print(i, 'O', a, '\n  G', lg[b, i])

if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    which = sys.argv[1:] or ['fn', 'solve']
    if 'fn' in which:
        for n in ['nmpc_tt', 't_trajectory', 'race_track_2']:
            fn_level(n)
    if 'solve' in which:
        solve_level('t_trajectory')
        solve_level('nmpc_tt')
        solve_level('race_track_2')
