"""GPU-side debugging aid (not a test): function-level and solve-level comparison with the CPU oracle."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
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


def golden_diff(name):
    """first diverging iteration between oracle and GPU on a golden fixture"""
    G = np.load(Path(__file__).resolve().parent.parent / 'tests' / 'golden' / f'solves_{name}.npz')
    sc = b200nmpc.SCENARIOS[name]
    lbx, ubx, lbg, ubg = sc.bounds()
    B = len(G['f'])
    s = b200nmpc.nlpsol('s', 'ipm', sc, max_batch=B)
    dbg = torch.zeros((B, 101, 8), dtype=torch.float64, device='cuda')
    b200nmpc._ffi.lib().nmpc_set_debug_log(s._h, dbg.data_ptr(), 101)
    s(x0=G['x0'], p=G['p'], lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg)
    st = s.stats(); lg = dbg.cpu().numpy()
    osp = oracle.make_spec(sc.T, sc.N, sc.n_obs); obs = sc.obstacle_table()
    print('golden', name, 'status O', G['status'], 'G', st['return_status'], 'iters O', G['iters'], 'G', st['iter_count'])
    for b in range(B):
        lo = oracle.solve_log(osp, obs, G['p'][b], G['x0'][b], lbx, ubx, lbg, ubg)['log']
        n = min(len(lo), int(st['iter_count'][b]))
        d = None
        for i in range(n):
            if not np.allclose(lo[i], lg[b, i], rtol=1e-6, atol=1e-12):
                d = i; break
        if d is not None or len(lo) != int(st['iter_count'][b]):
            print(f'-- instance {b}: first difference at iteration {d} (oracle {len(lo)} its, gpu {int(st["iter_count"][b])})')
            if d is not None:
                for i in range(max(0, d - 1), min(n, d + 3)):
                    print(i, 'O', lo[i], '\n  G', lg[b, i])


if __name__ == '__main__' and 'golden' in sys.argv[1:]:
    for n in ['nmpc_tt', 't_trajectory']:
        golden_diff(n)


def scaling_experiment():
    sc = b200nmpc.SCENARIOS['nmpc_tt']
    B = 64
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = b200nmpc.random_instances(sc, B, seed=2024 + 15)
    rng = np.random.default_rng(5)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N)) + 0.01 * rng.standard_normal((B, sc.n_w))
    osp = oracle.make_spec(sc.T, sc.N, sc.n_obs); obs = sc.obstacle_table()
    for scaling in (1, 2, 3, 0):
        ro = oracle.solve(osp, obs, p, x0, lbx, ubx, lbg, ubg, scaling=scaling)
        s = b200nmpc.nlpsol('s', 'ipm', sc, {'ipopt': {'nlp_scaling_method': 'gradient-based' if scaling else 'none', '_scaling_debug_mode': scaling}}, max_batch=B)
        dbg = torch.zeros((B, 101, 8), dtype=torch.float64, device='cuda')
        b200nmpc._ffi.lib().nmpc_set_debug_log(s._h, dbg.data_ptr(), 101)
        rg = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg); st = s.stats(); lg = dbg.cpu().numpy()
        print('scaling', scaling, 'oracle conv', (ro['status'] == 0).sum(), 'gpu conv', (st['return_status'] == 0).sum(),
              'status equal', (ro['status'] == st['return_status']).sum(), 'iters equal', (ro['iters'] == st['iter_count']).sum(), 'of', B)
        # minimum dual infeasibility reached in the last 5 logged iterations, oracle vs gpu, for instances the oracle converged
        worst = []
        for b in np.where(ro['status'] == 0)[0][:64]:
            lo = oracle.solve_log(osp, obs, p[b], x0[b], lbx, ubx, lbg, ubg, scaling=scaling)['log']
            n = int(st['iter_count'][b])
            worst.append((lo[-1, 3], lg[b, max(0, n - 1), 3], len(lo), n))
        w = np.array(worst)
        print('  last-logged inf_du: oracle median %.2e gpu median %.2e ; gpu>10x oracle in %d of %d' % (
            np.median(w[:, 0]), np.median(w[:, 1]), (w[:, 1] > 10 * w[:, 0]).sum(), len(w)))


if __name__ == '__main__' and 'scaling' in sys.argv[1:]:
    scaling_experiment()


def scaling_trace():
    sc = b200nmpc.SCENARIOS['nmpc_tt']
    B = 64
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = b200nmpc.random_instances(sc, B, seed=2024 + 15)
    rng = np.random.default_rng(5)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N)) + 0.01 * rng.standard_normal((B, sc.n_w))
    osp = oracle.make_spec(sc.T, sc.N, sc.n_obs); obs = sc.obstacle_table()
    scaling = 3
    ro = oracle.solve(osp, obs, p, x0, lbx, ubx, lbg, ubg, scaling=scaling)
    s = b200nmpc.nlpsol('s', 'ipm', sc, {'ipopt': {'_scaling_debug_mode': scaling}}, max_batch=B)
    dbg = torch.zeros((B, 101, 8), dtype=torch.float64, device='cuda')
    b200nmpc._ffi.lib().nmpc_set_debug_log(s._h, dbg.data_ptr(), 101)
    rg = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg); st = s.stats(); lg = dbg.cpu().numpy()
    bad = np.where((ro['status'] == 0) & (st['return_status'] != 0))[0][:2]
    for b in bad:
        lo = oracle.solve_log(osp, obs, p[b], x0[b], lbx, ubx, lbg, ubg, scaling=scaling)['log']
        n = int(st['iter_count'][b])
        print(f'-- instance {b}: oracle st {ro["status"][b]} it {len(lo)} | gpu st {st["return_status"][b]} it {n}')
        d = next((i for i in range(min(n, len(lo))) if not np.allclose(lo[i][:7], lg[b, i][:7], rtol=1e-5, atol=1e-12)), None)
        print('first diff', d)
        for i in range(max(0, (d or 0) - 2), min(max(n, len(lo)), (d or 0) + 12)):
            a = lo[i] if i < len(lo) else np.zeros(8)
            print(i, 'O', a, '\n  G', lg[b, i])
        for i in range(max(0, n - 14), n): print('TAIL', i, ' '.join('%.12e' % v for v in lg[b, i]))
        lam = rg['lam_g'][b].reshape(sc.N + 1, -1); lamo = ro['lam_g'][b].reshape(sc.N + 1, -1)
        print('max |lam_g| gpu per row type', np.abs(lam).max(axis=0), '\n oracle', np.abs(lamo).max(axis=0))


if __name__ == '__main__' and 'trace' in sys.argv[1:]:
    scaling_trace()


def scaling_resid():
    sc = b200nmpc.SCENARIOS['nmpc_tt']
    B = 64
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = b200nmpc.random_instances(sc, B, seed=2024 + 15)
    rng = np.random.default_rng(5)
    x0 = np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (B, sc.N)) + 0.01 * rng.standard_normal((B, sc.n_w))
    osp = oracle.make_spec(sc.T, sc.N, sc.n_obs); obs = sc.obstacle_table()
    scaling = 3
    ro = oracle.solve(osp, obs, p, x0, lbx, ubx, lbg, ubg, scaling=scaling)
    s = b200nmpc.nlpsol('s', 'ipm', sc, {'ipopt': {'_scaling_debug_mode': scaling}}, max_batch=B)
    rg = s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg); st = s.stats()
    b = int(np.where((ro['status'] == 0) & (st['return_status'] != 0))[0][0])
    np.set_printoptions(linewidth=250, precision=3)
    for name, r in (('GPU', rg), ('ORACLE', ro)):
        e = oracle.evaluate(osp, obs, r['x'][b], p[b])
        res = e['grad'] + e['J'].T @ r['lam_g'][b] + r['lam_x'][b]
        print(name, 'stationarity residual by stage (rows) x control (cols):\n', res.reshape(sc.N, 6))
    dl = (rg['lam_g'][b] - ro['lam_g'][b]).reshape(sc.N + 1, -1)
    print('lam_g gpu - oracle:\n', dl)
    print('lam_g oracle:\n', ro['lam_g'][b].reshape(sc.N + 1, -1))
    print('x gpu - oracle:\n', (rg['x'][b] - ro['x'][b]).reshape(sc.N, 6))
    print('lam_x gpu - oracle:\n', (rg['lam_x'][b] - ro['lam_x'][b]).reshape(sc.N, 6))
    g = rg['g'][b].reshape(sc.N + 1, -1)
    print('g gpu z-row:', g[:, 0], '\n theta row', g[:, 1])


if __name__ == '__main__' and 'resid' in sys.argv[1:]:
    scaling_resid()


def drift():
    """per-iteration relative deviation GPU vs oracle (f, inf_du, alpha_pr columns) for scaling modes 2 and 3"""
    sc = b200nmpc.SCENARIOS['nmpc_tt']
    B = 8
    lbx, ubx, lbg, ubg = sc.bounds()
    p, _ = b200nmpc.random_instances(sc, 64, seed=2024 + 15); p = p[:B]
    rng = np.random.default_rng(5)
    x0 = (np.tile(np.array([16.0, 0, 0, 0, 0, 0]), (64, sc.N)) + 0.01 * rng.standard_normal((64, sc.n_w)))[:B]
    osp = oracle.make_spec(sc.T, sc.N, sc.n_obs); obs = sc.obstacle_table()
    for scaling in (2, 3):
        s = b200nmpc.nlpsol('s', 'ipm', sc, {'ipopt': {'_scaling_debug_mode': scaling}}, max_batch=B)
        dbg = torch.zeros((B, 101, 8), dtype=torch.float64, device='cuda')
        b200nmpc._ffi.lib().nmpc_set_debug_log(s._h, dbg.data_ptr(), 101)
        s(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg); st = s.stats(); lg = dbg.cpu().numpy()
        for b in range(3):
            lo = oracle.solve_log(osp, obs, p[b], x0[b], lbx, ubx, lbg, ubg, scaling=scaling)['log']
            n = min(len(lo), int(st['iter_count'][b]))
            rel = [max(abs(lo[i][c] - lg[b, i][c]) / max(1e-300, abs(lo[i][c])) for c in (1, 3, 5)) for i in range(n)]
            print('mode', scaling, 'inst', b, 'its', len(lo), int(st['iter_count'][b]), ' '.join('%.0e' % r for r in rel))


if __name__ == '__main__' and 'drift' in sys.argv[1:]:
    drift()
