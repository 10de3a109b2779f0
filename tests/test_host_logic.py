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
