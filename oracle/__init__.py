"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/nmpc_oracle.cpp).

PARITY UNPINNED (see nmpc_oracle.cpp / nlp_ref.py headers): the reference has no golden
vectors and CasADi/IPOPT is not available, so the oracle is pinned only against
  * torch.autograd derivatives of the literal formulas (oracle/nlp_ref.py), and
  * KKT certificates evaluated independently in tests/, and
  * oracle/ipm_fullspace.py, a structurally independent restatement of the main loop (full-space KKT system, inertia
    counted from an LDL^T factorisation, autograd derivatives) that must reproduce the C++ oracle's iteration logs.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libnmpc_oracle.so"
_lib = None

STATUS_NAMES = {0: "Solve_Succeeded", 1: "Maximum_Iterations_Exceeded", 2: "Restoration_Failed",
                3: "Search_Direction_Becomes_Too_Small", 4: "Invalid_Number_Detected", 5: "Error_In_Step_Computation",
                6: "Infeasible_Problem_Detected"}
STAT_COLUMNS = ("factorizations", "soc_accepted", "resto_calls", "resto_iters", "watchdog_starts", "soft_resto_steps",
                "filter_resets", "slack_safeguards")


class OracleSpec(C.Structure):
    _fields_ = [("T", C.c_double), ("N", C.c_int32), ("n_obs", C.c_int32),
                ("w1", C.c_double), ("w2", C.c_double), ("vfov", C.c_double), ("hfov", C.c_double), ("model", C.c_int32)]

    # model 0: 8 states / 6 controls / p[11] / rows [z, theta, X5, X6, X7, obstacles]  (the Python scripts)
    # model 1: 5 states / 3 controls / p[8]  / rows [z, theta, obstacles]              (MATLAB/Dynamic Obstacles/NMPC_TT.m)
    @property
    def nu(self): return 3 if self.model else 6
    @property
    def npar(self): return 8 if self.model else 11
    @property
    def nw(self): return self.nu * self.N
    @property
    def ng(self): return ((2 if self.model else 5) + self.n_obs) * (self.N + 1)


def build(force: bool = False) -> Path:
    src = _HERE / "nmpc_oracle.cpp"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.oracle_eval.restype = C.c_int
        _lib.oracle_solve.restype = C.c_int
        _lib.oracle_solve_log.restype = C.c_int
        _lib.oracle_eval_traj.restype = C.c_int
        _lib.oracle_solve_traj.restype = C.c_int
        _lib.oracle_solve_warm.restype = C.c_int
        _lib.oracle_set_option.restype = C.c_int
        _lib.oracle_set_option.argtypes = [C.c_char_p, C.c_double]
    return _lib


def set_option(name: str, value: float):
    """Override one algorithmic option of the oracle (Options in nmpc_oracle.cpp), e.g. set_option("resto", 0)."""
    lib().oracle_set_option(name.encode(), float(value))


def clear_options():
    lib().oracle_clear_options()


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_spec(T, N, n_obs, w1=1.0, w2=2.0, vfov=1.0, hfov=1.0, model=0) -> OracleSpec:
    return OracleSpec(float(T), int(N), int(n_obs), float(w1), float(w2), float(vfov), float(hfov), int(model))


def obstacle_table(obstacles, uav_r=5.0) -> np.ndarray:
    """[(cx,cy,r_obs)] -> [n_obs][3] = cx, cy, uav_r + r_obs (the constant of NMPC_TT.py:241)."""
    o = np.asarray(obstacles, dtype=np.float64).reshape(-1, 3).copy()
    o[:, 2] += uav_r
    return o


def evaluate(spec: OracleSpec, obs: np.ndarray, w, p, lam_g=None, sigma=1.0, hessian=False, target_traj=None):
    N, n_obs = spec.N, spec.n_obs
    nw, ng = spec.nw, spec.ng
    w, p, obs = _c(w), _c(p), _c(obs)
    f = C.c_double()
    g = np.zeros(ng); grad = np.zeros(nw); J = np.zeros((ng, nw)); X = np.zeros((N + 1, 8))
    H = np.zeros((nw, nw)) if hessian else None
    lam = _c(lam_g) if lam_g is not None else np.zeros(ng)
    tg = None if target_traj is None else _c(target_traj).reshape(N, 2)     # per-stage predicted target
    lib().oracle_eval_traj(C.byref(spec), _dp(obs), _dp(w), _dp(p), _dp(tg), C.c_double(sigma), _dp(lam),
                           C.byref(f), _dp(g), _dp(grad), _dp(J), _dp(H), _dp(X))
    return dict(f=f.value, g=g, grad=grad, J=J, H=H, X=X)


def solve(spec: OracleSpec, obs, p, x0, lbx, ubx, lbg, ubg, *, obs_per_instance=False, scaling=True,
          max_iter=0, tol=0.0, nthreads=None, want_g=True, want_lam=True, target_traj=None, lam_x0=None, lam_g0=None):
    """Batch solve: p (B,11), x0 (B,6N) -> dict(x,f,g,lam_x,lam_g,status,iters,stats).
    lam_x0 (B,6N), lam_g0 (B,n_g): multiplier guesses of the NON-REFERENCE warm-start mode (Options::ws_* in
    nmpc_oracle.cpp); a NaN in lam_x0[b, 0] cold-starts instance b."""
    N, n_obs = spec.N, spec.n_obs
    nw, ng = spec.nw, spec.ng
    p = _c(p).reshape(-1, spec.npar); B = p.shape[0]
    x0 = _c(x0).reshape(B, nw)
    lbx, ubx, lbg, ubg, obs = _c(lbx), _c(ubx), _c(lbg), _c(ubg), _c(obs)
    assert lbx.size == nw and ubx.size == nw and lbg.size == ng and ubg.size == ng
    assert obs.size == (B if obs_per_instance else 1) * n_obs * 3
    x = np.zeros((B, nw)); f = np.zeros(B)
    g = np.zeros((B, ng)) if want_g else None
    lam_x = np.zeros((B, nw)) if want_lam else None
    lam_g = np.zeros((B, ng)) if want_lam else None
    status = np.zeros(B, dtype=np.int32); iters = np.zeros(B, dtype=np.int32); stats = np.zeros((B, 8), dtype=np.int32)
    if nthreads is None:
        nthreads = min(B, os.cpu_count() or 1)
    tg = None if target_traj is None else _c(target_traj).reshape(B, N, 2)
    lx0 = None if lam_x0 is None else _c(lam_x0).reshape(B, nw)
    lg0 = None if lam_g0 is None else _c(lam_g0).reshape(B, ng)
    assert (lx0 is None) == (lg0 is None)
    lib().oracle_solve_warm(C.byref(spec), C.c_int(B), _dp(p), _dp(x0), _dp(tg), _dp(lx0), _dp(lg0), _dp(lbx), _dp(ubx), _dp(lbg), _dp(ubg),
                       _dp(obs), C.c_int(int(obs_per_instance)), C.c_int(int(scaling)), C.c_int(int(max_iter)),
                       C.c_double(float(tol)),
                       _dp(x), _dp(f), _dp(g), _dp(lam_x), _dp(lam_g), _ip(status), _ip(iters), _ip(stats),
                       C.c_int(int(nthreads)))
    return dict(x=x, f=f, g=g, lam_x=lam_x, lam_g=lam_g, status=status, iters=iters, stats=stats)


def solve_log(spec: OracleSpec, obs, p, x0, lbx, ubx, lbg, ubg, scaling=True, max_log=128):
    N = spec.N; nw = spec.nw
    x = np.zeros(nw); f = C.c_double(); st = C.c_int32(); it = C.c_int32()
    log = np.zeros((max_log, 9))
    n = lib().oracle_solve_log(C.byref(spec), _dp(_c(p)), _dp(_c(x0)), _dp(_c(lbx)), _dp(_c(ubx)), _dp(_c(lbg)),
                               _dp(_c(ubg)), _dp(_c(obs)), C.c_int(int(scaling)), _dp(x), C.byref(f), C.byref(st),
                               C.byref(it), _dp(log), C.c_int(max_log))
    return dict(x=x, f=f.value, status=st.value, iters=it.value, log=log[:min(n, max_log)])
