"""TEST INFRASTRUCTURE ONLY (oracle) -- literal restatement of the reference NLP.

PARITY UNPINNED: the reference (devsonni/MPC-Implementation) ships no tests, golden
vectors or recorded outputs for this path, and CasADi/IPOPT cannot be imported in this
image, so nothing here was checked against the reference's own numbers.  This file is a
line-by-line restatement of the symbolic problem the reference hands to CasADi:

  dynamics / Euler rollout ...... Python/NMPC_TT.py:139-148, :160-167
  objective (distance + FOV) .... Python/NMPC_TT.py:193-221  (w1=1, w2=2, VFOV=HFOV=1)
  constraint rows g .............. Python/NMPC_TT.py:234-244  (Race Track 2.py:247-264 for 10 obstacles)
  bounds ......................... Python/NMPC_TT.py:62-89, :269-306

It is written in torch.float64 so that torch.autograd supplies gradient / Jacobian /
Hessian values that are *mechanically* derived from the literal formulas -- the same role
CasADi's AD plays in the reference.  The C++ oracle (nmpc_oracle.cpp) and the CUDA kernels
are both checked against it in tests/.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np
import torch

NX = 8   # states  [x, y, z, theta, psi, phi_g, shi_g, theta_g]   NMPC_TT.py:105-114
NU = 6   # controls [v, w2u, w3u, w1g, w2g, w3g]                  NMPC_TT.py:128-135
NP = 11  # p = [state(8); x_t; y_t; theta_t]                       NMPC_TT.py:154, :350-353

PI = math.pi


@dataclass
class RefSpec:
    """Constants that the reference edits in source (one script = one RefSpec)."""
    T: float = 1.0
    N: int = 15
    obstacles: Sequence[Tuple[float, float, float]] = ((175.0, 820.0, 30.0), (-134.0, 155.0, 30.0), (441.0, 343.0, 30.0))
    uav_r: float = 5.0
    w1: float = 1.0
    w2: float = 2.0
    vfov: float = 1.0
    hfov: float = 1.0

    @property
    def n_obs(self) -> int:
        return len(self.obstacles)

    @property
    def n_w(self) -> int:
        return NU * self.N

    @property
    def n_g(self) -> int:
        return (5 + self.n_obs) * (self.N + 1)


def rollout(spec: RefSpec, w: torch.Tensor, p: torch.Tensor) -> List[torch.Tensor]:
    """X[:,0]=P[0:8]; X[:,k+1]=X[:,k]+T*f_u(X[:,k],U[:,k])  (NMPC_TT.py:160-167).

    w is vec(U) column-major: w[6k+i] = U[i,k] (NMPC_TT.py:247-248)."""
    X = [p[0:8]]
    for k in range(spec.N):
        st = X[k]
        con = w[NU * k: NU * (k + 1)]
        v_u, om2, om3, om1g, om2g, om3g = con
        theta_u, psi_u = st[3], st[4]
        rhs = torch.stack([
            v_u * torch.cos(psi_u) * torch.cos(theta_u),
            v_u * torch.sin(psi_u) * torch.cos(theta_u),
            v_u * torch.sin(theta_u),
            om2, om3, om1g, om2g, om3g])                      # NMPC_TT.py:139-148
        X.append(st + spec.T * rhs)
    return X


def objective(spec: RefSpec, w: torch.Tensor, p: torch.Tensor, target_traj: torch.Tensor | None = None) -> torch.Tensor:
    """Literal transcription of NMPC_TT.py:193-221.  target_traj [N, 2] (optional) replaces the constant
    target (p[8], p[9]) by a per-stage prediction; the reference itself keeps the target fixed over the horizon."""
    X = rollout(spec, w, p)
    obj = torch.zeros((), dtype=w.dtype)
    VFOV, HFOV = spec.vfov, spec.hfov
    for k in range(spec.N):
        st = X[k]
        a = (st[2] * torch.tan(st[6] + VFOV / 2) - st[2] * torch.tan(st[6] - VFOV / 2)) / 2
        b = (st[2] * torch.tan(st[5] + HFOV / 2) - st[2] * torch.tan(st[5] - HFOV / 2)) / 2
        A = (torch.cos(st[7])) ** 2 / a ** 2 + (torch.sin(st[7])) ** 2 / b ** 2
        B = 2 * torch.cos(st[7]) * torch.sin(st[7]) * ((1 / a ** 2) - (1 / b ** 2))
        C = (torch.sin(st[7])) ** 2 / a ** 2 + (torch.cos(st[7])) ** 2 / b ** 2
        X_E = st[0] + a + st[2] * torch.tan(st[6] - VFOV / 2)
        Y_E = st[1] + b + st[2] * torch.tan(st[5] - HFOV / 2)
        xt, yt = (p[8], p[9]) if target_traj is None else (target_traj[k, 0], target_traj[k, 1])
        obj = obj + spec.w1 * torch.sqrt((st[0] - xt) ** 2 + (st[1] - yt) ** 2) + \
            spec.w2 * ((A * (xt - X_E) ** 2 + B * (yt - Y_E) * (xt - X_E) + C * (yt - Y_E) ** 2) - 1)
    return obj


def constraints(spec: RefSpec, w: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """Literal transcription of NMPC_TT.py:234-244: rows [z,theta,X5,X6,X7,obs_1..] per stage 0..N."""
    X = rollout(spec, w, p)
    rows = []
    for k in range(spec.N + 1):
        st = X[k]
        rows += [st[2], st[3], st[5], st[6], st[7]]
        for (cx, cy, r) in spec.obstacles:
            rows.append(-torch.sqrt((st[0] - cx) ** 2 + (st[1] - cy) ** 2) + (spec.uav_r + r))
    return torch.stack(rows)


def bounds(spec: RefSpec):
    """lbx, ubx, lbg, ubg as numpy float64 (NMPC_TT.py:62-89, :269-306)."""
    N = spec.N
    lbx = np.zeros(NU * N)
    ubx = np.zeros(NU * N)
    lo = [14.0, -PI / 30, -PI / 21, -PI / 30, -PI / 30, -PI / 30]
    hi = [30.0, PI / 30, PI / 21, PI / 30, PI / 30, PI / 30]
    for i in range(NU):
        lbx[i::NU] = lo[i]
        ubx[i::NU] = hi[i]
    rows = 5 + spec.n_obs
    lbg = np.zeros(rows * (N + 1))
    ubg = np.zeros(rows * (N + 1))
    glo = [75.0, -0.2618, -PI / 6, -PI / 6, -PI / 2]
    ghi = [150.0, 0.2618, PI / 6, PI / 6, PI / 2]
    for i in range(5):
        lbg[i::rows] = glo[i]
        ubg[i::rows] = ghi[i]
    for j in range(spec.n_obs):
        lbg[5 + j::rows] = -np.inf
        ubg[5 + j::rows] = 0.0
    return lbx, ubx, lbg, ubg


def eval_all(spec: RefSpec, w: np.ndarray, p: np.ndarray, lam_g: np.ndarray | None = None, sigma: float = 1.0,
             target_traj: np.ndarray | None = None):
    """f, g, grad f, J, and (if lam_g given) Hessian of sigma*f + lam_g^T g -- all by autograd."""
    wt = torch.tensor(np.asarray(w, dtype=np.float64), requires_grad=True)
    pt = torch.tensor(np.asarray(p, dtype=np.float64))
    tt = None if target_traj is None else torch.tensor(np.asarray(target_traj, dtype=np.float64)).reshape(spec.N, 2)
    f = objective(spec, wt, pt, tt)
    g = constraints(spec, wt, pt)
    grad = torch.autograd.grad(f, wt, retain_graph=True)[0]
    J = torch.autograd.functional.jacobian(lambda ww: constraints(spec, ww, pt), wt)
    out = dict(f=float(f.detach()), g=g.detach().numpy(), grad=grad.numpy(), J=J.numpy())
    if lam_g is not None:
        lt = torch.tensor(np.asarray(lam_g, dtype=np.float64))
        H = torch.autograd.functional.hessian(
            lambda ww: sigma * objective(spec, ww, pt, tt) + (lt * constraints(spec, ww, pt)).sum(), wt)
        out["H"] = H.numpy()
    return out


# ----------------------------------------------------------------------------------------------
# plant / target step of the closed loop (NMPC_TT.py:13-30) and FOV-centre bookkeeping (:399-402)
# ----------------------------------------------------------------------------------------------
def f_u(x: np.ndarray, u: np.ndarray) -> np.ndarray:
    v, om2, om3, o1, o2, o3 = u
    th, ps = x[3], x[4]
    return np.array([v * math.cos(ps) * math.cos(th), v * math.sin(ps) * math.cos(th), v * math.sin(th),
                     om2, om3, o1, o2, o3])


def shift_timestep(T: float, x0: np.ndarray, u: np.ndarray, xs: np.ndarray, con_t: Tuple[float, float]):
    """u is (6,N).  Returns x0+, u0+ (warm start: drop first column, repeat last), xs+."""
    x0n = x0 + T * f_u(x0, u[:, 0])
    u0 = np.concatenate([u[:, 1:], u[:, -1:]], axis=1)
    f_t = np.array([con_t[0] * math.cos(xs[2]), con_t[0] * math.sin(xs[2]), con_t[1]])
    xsn = xs + T * f_t
    return x0n, u0, xsn


def fov_centre(x0: np.ndarray, vfov: float = 1.0, hfov: float = 1.0):
    a_p = (x0[2] * math.tan(x0[6] + vfov / 2) - x0[2] * math.tan(x0[6] - vfov / 2)) / 2
    b_p = (x0[2] * math.tan(x0[5] + hfov / 2) - x0[2] * math.tan(x0[5] - hfov / 2)) / 2
    return (x0[0] + a_p + x0[2] * math.tan(x0[6] - vfov / 2),
            x0[1] + b_p + x0[2] * math.tan(x0[5] - hfov / 2))


# ----------------------------------------------------------------------------------------------
# The gimbal-less tracker of MATLAB/Dynamic Obstacles/NMPC_TT.m (model 1): 5 states, 3 controls, p = [state(5); target(3)],
# distance-only cost, rows [z, theta] per stage.  Literal restatement, independent of the 8-state code above.
# ----------------------------------------------------------------------------------------------
NX5, NU5, NP5 = 5, 3, 8     # NMPC_TT.m:25-35, :37


@dataclass
class RefSpec5:
    T: float = 0.2          # NMPC_TT.m:9
    N: int = 15             # NMPC_TT.m:10

    @property
    def n_w(self) -> int:
        return NU5 * self.N

    @property
    def n_g(self) -> int:
        return 2 * (self.N + 1)


def rollout5(spec: RefSpec5, w: torch.Tensor, p: torch.Tensor) -> List[torch.Tensor]:
    """X(:,1) = P(1:5); X(:,k+1) = X(:,k) + T*f_u(X(:,k), U(:,k))   NMPC_TT.m:45-51, rhs :33-34"""
    X = [p[0:5]]
    for k in range(spec.N):
        st = X[k]
        v_u, om2, om3 = w[NU5 * k: NU5 * (k + 1)]
        th, ps = st[3], st[4]
        rhs = torch.stack([v_u * torch.cos(ps) * torch.cos(th), v_u * torch.sin(ps) * torch.cos(th), v_u * torch.sin(th), om2, om3])
        X.append(st + spec.T * rhs)
    return X


def objective5(spec: RefSpec5, w: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """obj = sum_{k=1..N} sqrt((X(1,k)-P(6))^2 + (X(2,k)-P(7))^2)   NMPC_TT.m:100-104 (1-based: stages 0..N-1)"""
    X = rollout5(spec, w, p)
    obj = torch.zeros((), dtype=w.dtype)
    for k in range(spec.N):
        obj = obj + torch.sqrt((X[k][0] - p[5]) ** 2 + (X[k][1] - p[6]) ** 2)
    return obj


def constraints5(spec: RefSpec5, w: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """g = [X(3,k); X(4,k)] for k = 1..N+1   NMPC_TT.m:107-111"""
    X = rollout5(spec, w, p)
    rows = []
    for k in range(spec.N + 1):
        rows += [X[k][2], X[k][3]]
    return torch.stack(rows)


def bounds5(spec: RefSpec5):
    """NMPC_TT.m:14-22, :127-133"""
    N = spec.N
    lbx = np.tile(np.array([14.0, -PI / 30, -PI / 21]), N)
    ubx = np.tile(np.array([30.0, PI / 30, PI / 21]), N)
    lbg = np.tile(np.array([75.0, -0.2618]), N + 1)
    ubg = np.tile(np.array([150.0, 0.2618]), N + 1)
    return lbx, ubx, lbg, ubg


def eval_all5(spec: RefSpec5, w: np.ndarray, p: np.ndarray, lam_g: np.ndarray | None = None, sigma: float = 1.0):
    wt = torch.tensor(np.asarray(w, dtype=np.float64), requires_grad=True)
    pt = torch.tensor(np.asarray(p, dtype=np.float64))
    f = objective5(spec, wt, pt)
    g = constraints5(spec, wt, pt)
    grad = torch.autograd.grad(f, wt, retain_graph=True)[0]
    J = torch.autograd.functional.jacobian(lambda ww: constraints5(spec, ww, pt), wt)
    out = dict(f=float(f.detach()), g=g.detach().numpy(), grad=grad.numpy(), J=J.numpy())
    if lam_g is not None:
        lt = torch.tensor(np.asarray(lam_g, dtype=np.float64))
        H = torch.autograd.functional.hessian(lambda ww: sigma * objective5(spec, ww, pt) + (lt * constraints5(spec, ww, pt)).sum(), wt)
        out["H"] = H.numpy()
    return out


def shift1(T: float, x0: np.ndarray, u: np.ndarray, xs: np.ndarray, con_t: Tuple[float, float] = (15.0, 0.12)):
    """MATLAB/Dynamic Obstacles/shift1.m: u is (N,3) (rows = stages).  Returns x0+, u0+ (drop first row, repeat last), xs+."""
    v, om2, om3 = u[0]
    th, ps = x0[3], x0[4]
    x0n = x0 + T * np.array([v * math.cos(ps) * math.cos(th), v * math.sin(ps) * math.cos(th), v * math.sin(th), om2, om3])
    xsn = xs + T * np.array([con_t[0] * math.cos(xs[2]), con_t[0] * math.sin(xs[2]), con_t[1]])
    u0 = np.concatenate([u[1:], u[-1:]], axis=0)
    return x0n, u0, xsn
