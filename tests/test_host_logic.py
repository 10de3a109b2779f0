"""CPU: scenario registry vs the reference scripts' constants, the C-ABI exports, host-side sharding."""
import ctypes
import math
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import nlp_ref

ROOT = Path(__file__).resolve().parent.parent


def test_schedules_match_reference_breakpoints(pkg):
    S = pkg.SCENARIOS
    pi = math.pi
    assert S["nmpc_tt"].schedule(0) == (12.0, 0.01) and S["nmpc_tt"].schedule(699) == (12.0, 0.01)     # NMPC_TT.py:25
    t = S["t_trajectory"].schedule                                                                        # T_Trajectory.py:25-57
    assert t(99) == (13.5, 0.0) and t(100) == (13.5, (pi / 2) / 12) and t(160)[1] == 0.0 and t(260)[1] == -(pi / 2) / 12
    assert t(1572)[1] == 0.0 and t(1573)[1] == (pi / 2) / 12
    pl = S["plus_trajectory"].schedule                                                                    # Plus Trajectory.py:25-69
    assert pl(100)[1] == 0.0 and pl(101) == (20.0, (pi / 2) * 5) and pl(102)[1] == 0.0 and pl(203)[1] == -(pi / 2) * 5
    r1 = S["race_trajectory_1"].schedule                                                                  # Race Trajectory 1.py:27-57
    assert r1(300) == (14.0, -(pi / 2) / 24) and r1(570)[1] == ((11 * pi) / 18) / 12 and r1(1535)[1] == (pi / 2) / 12
    r2 = S["race_track_2"].schedule                                                                       # Race Track 2.py:28-36
    assert r2(499)[1] == 0.0 and r2(500) == (12.0, pi / 100) and r2(1000)[1] == 0.0 and r2(1999)[1] == pi / 100
    assert S["10_obstacles"].schedule(300) == (13.0, -(pi / 2) / 24)                                      # 10_obstacles.py:28-32


def test_scenario_constants(pkg):
    S = pkg.SCENARIOS
    assert S["nmpc_tt"].T == 1.0 and S["nmpc_tt"].steps == 700 and S["nmpc_tt"].x_init[0] == 90.0         # NMPC_TT.py:57,:321,:339
    assert [s.steps for s in (S["t_trajectory"], S["plus_trajectory"], S["race_trajectory_1"], S["race_track_2"], S["10_obstacles"])] \
        == [1633, 1223, 1595, 2000, 1595]
    assert S["race_track_2"].n_g == 240 and S["nmpc_tt"].n_g == 128 and S["nmpc_tt"].n_w == 90
    assert np.allclose(S["race_track_2"].obstacle_table()[4], [1765, 550, 55])                             # Race Track 2.py:231-232,:243-244
    assert np.allclose(S["10_obstacles"].obstacle_table()[0], [500, 20, 105])
    for name, sc in S.items():
        rs = nlp_ref.RefSpec(T=sc.T, N=sc.N, obstacles=sc.obstacles)
        for a, b in zip(sc.bounds(), nlp_ref.bounds5(nlp_ref.RefSpec5(sc.T, sc.N)) if sc.model else nlp_ref.bounds(rs)):
            assert np.array_equal(a, b), name


def test_abi_exports_every_declared_symbol(pkg):
    hdr = (ROOT / "include" / "nmpc_b200.h").read_text()
    declared = set(re.findall(r"\b(nmpc_[a-z_0-9]+)\s*\(", hdr))
    assert {"nmpc_create", "nmpc_solve", "nmpc_solve_host", "nmpc_eval", "nmpc_step", "nmpc_destroy"} <= declared
    lib = ctypes.CDLL(str(pkg._ffi.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nmpc_b200.h but not exported"
    assert set(pkg._ffi.EXPORTS) == declared
    assert b"sm_100a" in pkg._ffi.lib().nmpc_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(pkg):
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pkg.nlpsol("solver", "ipm", pkg.SCENARIOS["nmpc_tt"])
    with pytest.raises(ValueError):
        pkg.nlpsol("solver", "sqpmethod", pkg.SCENARIOS["nmpc_tt"])


def test_spec_struct_layout_matches_header(pkg):
    """ctypes mirror of struct nmpc_spec has the field order / size the C header implies."""
    f = [n for n, _ in pkg._ffi.NmpcSpec._fields_]
    assert f == ["T", "N", "n_obs", "w1", "w2", "vfov", "hfov", "max_iter", "scaling", "tol", "max_batch", "fill", "model"]
    assert ctypes.sizeof(pkg._ffi.NmpcSpec) == 80


def test_random_instances_are_seeded_and_feasible(pkg):
    sc = pkg.SCENARIOS["nmpc_tt"]
    p1, vw1 = pkg.random_instances(sc, 64, 5)
    p2, vw2 = pkg.random_instances(sc, 64, 5)
    assert np.array_equal(p1, p2) and np.array_equal(vw1, vw2)
    _, _, lbg, ubg = sc.bounds()
    rows = np.stack([p1[:, 2], p1[:, 3], p1[:, 5], p1[:, 6], p1[:, 7]], axis=1)
    assert np.all(rows > lbg[:5]) and np.all(rows < ubg[:5])


def test_shard_range_partitions():
    from mpc_implementation_b200.sharding import shard_range
    for B in (0, 1, 7, 4096, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_result_dict_computes_lam_p_on_first_access():
    """sol['lam_p'] (CasADi's sixth result key, never read by the reference) costs nothing unless somebody reads it."""
    from mpc_implementation_b200.nlpsol import _Sol
    calls = []
    s = _Sol(dict(x=1, f=2), lambda: calls.append(1) or "LP")
    assert "lam_p" in s and set(s.keys()) == {"x", "f", "lam_p"} and not calls
    assert s["x"] == 1 and not calls
    assert s["lam_p"] == "LP" and s["lam_p"] == "LP" and len(calls) == 1
    assert dict(s.items())["lam_p"] == "LP" and s.get("lam_p") == "LP" and s.get("nope", 7) == 7
    t = _Sol(dict(x=1), None)
    assert "lam_p" not in t


def test_casadi_recorder_mechanics(pkg, oracle_mod, tmp_path, monkeypatch):
    """bench/run_casadi.py (the pinning hook) end to end WITHOUT CasADi: a stand-in `casadi` module whose nlpsol is backed
    by the CPU oracle, and a miniature script with the reference's call shape (NMPC_TT.py:267, :358-365 -- kwargs of
    DM columns, result dict of DMs, solver.stats()).  Checks what can be checked here: the wrapper records every call and
    unwraps itself, plotting imports are stubbed, --max-steps stops the script's loop, and the file it writes has the
    layout tests/test_oracle_solve.py::test_oracle_against_casadi_records and tests/test_gpu_parity.py::test_casadi_records
    load (so those two become live the day real records exist).  It does NOT pin anything: the numbers are the oracle's."""
    import importlib.util
    import sys
    import types
    sc = pkg.SCENARIOS["nmpc_tt"]
    sp = oracle_mod.make_spec(sc.T, sc.N, sc.n_obs, sc.w1, sc.w2, sc.vfov, sc.hfov)
    lbx, ubx, lbg, ubg = sc.bounds()

    class DM:                                                         # the sliver of casadi.DM the recorder and the script use
        def __init__(self, v):
            self.a = np.asarray(v.a if isinstance(v, DM) else v, dtype=np.float64).reshape(-1, 1)

        def full(self):
            return self.a

        def __float__(self):
            return float(self.a.reshape(-1)[0])

    calls = []

    def nlpsol(name, plugin, prob, opts):
        assert (name, plugin) == ("solver", "ipopt")

        class Fn:
            def __call__(self, *, x0, lbx, ubx, lbg, ubg, p):
                v = lambda d: DM(d).full().reshape(-1)
                r = oracle_mod.solve(sp, sc.obstacle_table(), v(p)[None], v(x0)[None], v(lbx), v(ubx), v(lbg), v(ubg))
                self._st = dict(return_status=oracle_mod.STATUS_NAMES[int(r["status"][0])], iter_count=int(r["iters"][0]), success=bool(r["status"][0] == 0))
                calls.append(1)
                return {k: DM(r[k][0]) for k in ("x", "f", "g", "lam_x", "lam_g")}

            def stats(self):
                return self._st
        return Fn()

    fake = types.ModuleType("casadi")
    fake.DM, fake.nlpsol, fake.__version__ = DM, nlpsol, "stand-in"
    monkeypatch.setitem(sys.modules, "casadi", fake)
    script = tmp_path / "Python" / "NMPC_TT.py"
    script.parent.mkdir()
    p0 = list(sc.x_init) + list(sc.target_init)
    script.write_text(
        "import casadi as ca\nimport numpy as np\nimport matplotlib.pyplot as plt\nfrom mayavi import mlab\n"
        "import sys; sys.path.insert(0, %r)\nimport b200nmpc\nsc = b200nmpc.SCENARIOS['nmpc_tt']\n"
        "lbx, ubx, lbg, ubg = sc.bounds()\nsolver = ca.nlpsol('solver', 'ipopt', {}, {})\n"
        "p = np.array(%r); u0 = np.zeros(sc.n_w)\n"
        "if __name__ == '__main__':\n"
        "    for mpc_iter in range(50):\n"
        "        sol = solver(x0=ca.DM(u0), lbx=ca.DM(lbx), ubx=ca.DM(ubx), lbg=ca.DM(lbg), ubg=ca.DM(ubg), p=ca.DM(p))\n"
        "        u0 = np.roll(sol['x'].full().reshape(-1), -6)\n"
        "    plt.plot([0], [0]); mlab.show()\n" % (str(ROOT), p0))
    spec = importlib.util.spec_from_file_location("run_casadi", ROOT / "bench" / "run_casadi.py")
    rc = importlib.util.module_from_spec(spec); spec.loader.exec_module(rc)
    monkeypatch.setattr(sys, "argv", ["run_casadi.py", "--reference-root", str(tmp_path), "--scripts", "NMPC_TT.py", "--max-steps", "3",
                                      "--out", str(tmp_path / "golden")])
    before = set(sys.modules)
    try:
        assert rc.main() == 0
    finally:
        for k in set(sys.modules) - before:            # the recorder's inert plotting stubs must not outlive this test
            if isinstance(sys.modules[k], rc._Stub):
                del sys.modules[k]
    assert len(calls) == 3 and fake.nlpsol is nlpsol                     # stopped after --max-steps; casadi.nlpsol restored
    C = np.load(tmp_path / "golden" / "casadi_nmpc_tt.npz", allow_pickle=False)
    assert C["p"].shape == (3, 11) and C["x0"].shape == (3, sc.n_w) and C["x"].shape == (3, sc.n_w) and C["g"].shape == (3, sc.n_g)
    assert C["lam_x"].shape == (3, sc.n_w) and C["lam_g"].shape == (3, sc.n_g) and C["f"].shape == (3,)
    assert C["status"].dtype == np.int32 and C["iters"].dtype == np.int32 and str(C["casadi_version"]) == "stand-in"
    assert np.array_equal(C["p"][0], np.array(p0)) and not C["x0"][0].any() and np.array_equal(C["x0"][1], np.roll(C["x"][0], -6))
    # ... and the loader side: the recorded file drives the same comparison the pinning tests run
    r = oracle_mod.solve(sp, sc.obstacle_table(), C["p"], C["x0"], lbx, ubx, lbg, ubg)
    assert np.array_equal(r["status"], C["status"]) and np.array_equal(r["iters"], C["iters"]) and np.allclose(r["f"], C["f"], rtol=1e-12)


def test_plain_c_caller_builds_against_the_header(pkg, tmp_path):
    """include/nmpc_b200.h is a C header (not only C++) and examples/c_abi_demo.c -- a plain-C closed loop over
    nmpc_solve_host, no Python / torch / CUDA headers -- compiles and links against libnmpc_b200.so.  Without a GPU the
    program must fail LOUDLY in nmpc_create (no CPU fallback); with one it prints the known answer of T_Trajectory.py's
    first solve (its B200 output is committed as profiles/r2/c_abi_demo.log)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = tmp_path / "c_abi_demo"
    libdir = pkg._ffi.LIB_PATH.parent
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-std=c11", "-I", str(ROOT / "include"), "-o", str(exe), str(ROOT / "examples" / "c_abi_demo.c"),
                           "-L", str(libdir), "-lnmpc_b200", f"-Wl,-rpath,{libdir}", "-lm"])
    r = subprocess.run([str(exe), "1"], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "f = 248.10109322" in r.stdout, (r.stdout, r.stderr)
    else:
        assert r.returncode == 2 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)
