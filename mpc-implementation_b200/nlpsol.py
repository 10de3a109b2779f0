"""CasADi-shaped facade over the C ABI: the one line a reference script swaps.

Reference (Python/NMPC_TT.py:250-267, :358-367):
    solver = ca.nlpsol('solver', 'ipopt', nlp_prob, opts)
    sol = solver(x0=args['x0'], lbx=args['lbx'], ubx=args['ubx'], lbg=args['lbg'], ubg=args['ubg'], p=args['p'])
    u = ca.reshape(sol['x'], n_controls, N)
Here:
    solver = b200nmpc.nlpsol('solver', 'ipm', scenario_or_dict, opts)       # nlp_prob -> problem constants
    sol = solver(x0=..., lbx=..., ubx=..., lbg=..., ubg=..., p=...)          # same keywords, same layouts
    u = sol['x'].reshape(N, 6).T

Every argument may be a single instance ((n,), (n,1), list) or a batch (B, n); host containers (numpy, lists)
go through nmpc_solve_host (copies inside), torch CUDA float64 tensors are used in place on torch's current
stream.  Like CasADi's nlpsol the call never raises on a non-converged instance: the outcome is in
solver.stats() ('return_status', 'success', 'iter_count'), which the reference never reads (:358-367).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional

import numpy as np

from . import _ffi
from .scenarios import NP, NU, Scenario

try:  # torch is plumbing for device memory / streams only
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_cuda_tensor(a) -> bool:
    return torch is not None and isinstance(a, torch.Tensor) and a.is_cuda


def _spec_from(problem, opts: Optional[dict]) -> Dict[str, Any]:
    if isinstance(problem, Scenario):
        d = dict(T=problem.T, N=problem.N, obstacles=problem.obstacle_table(), w1=problem.w1, w2=problem.w2,
                 vfov=problem.vfov, hfov=problem.hfov, model=problem.model)
    else:
        d = dict(problem)
        if "obstacles" in d:
            o = np.asarray(d["obstacles"], dtype=np.float64).reshape(-1, 3).copy()
            o[:, 2] += float(d.get("uav_r", 5.0))
            d["obstacles"] = o
        else:
            d["obstacles"] = np.zeros((int(d.get("n_obs", 0)), 3))
    ip = dict((opts or {}).get("ipopt", {}))
    d["max_iter"] = int(ip.get("max_iter", 100))          # NMPC_TT.py:259
    d["tol"] = float(ip.get("tol", 1e-8))
    d["scaling"] = 0 if ip.get("nlp_scaling_method", "gradient-based") == "none" else int(ip.get("_scaling_debug_mode", 1))
    # NON-REFERENCE mode (no script sets it): IPOPT's warm_start_init_point and its pushes; lam_x0 / lam_g0 are ignored without it
    d["warm"] = str(ip.get("warm_start_init_point", "no")).lower() == "yes"
    d["warm_opts"] = (float(ip.get("mu_init", 0.0)) if d["warm"] else 0.0, float(ip.get("warm_start_bound_push", 0.0)),
                      float(ip.get("warm_start_bound_frac", 0.0)), float(ip.get("warm_start_slack_bound_push", 0.0)),
                      float(ip.get("warm_start_slack_bound_frac", 0.0)), float(ip.get("warm_start_mult_bound_push", 0.0)))
    return d


class _Sol(dict):
    """Result dict of a solve.  sol['lam_p'] -- the sixth key of CasADi's result dict, which no reference script reads -- is
    computed on first access (one small launch, nmpc_lam_p) instead of on every call."""

    def __init__(self, items, lam_p_fn=None):
        super().__init__(items)
        self._lam_p_fn = lam_p_fn
        if lam_p_fn is not None:
            super().__setitem__("lam_p", None)

    def __getitem__(self, k):
        if k == "lam_p" and self._lam_p_fn is not None:
            super().__setitem__("lam_p", self._lam_p_fn())
            self._lam_p_fn = None
        return super().__getitem__(k)

    def get(self, k, default=None):
        return self[k] if k in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]


class Solver:
    """Callable returned by nlpsol(); owns one nmpc_handle."""

    def __init__(self, name: str, problem, opts: Optional[dict] = None, device: int = 0, max_batch: int = 1, fill: int = 1):
        self.name = name
        self.fill = int(fill)
        d = _spec_from(problem, opts)
        self.T, self.N = float(d["T"]), int(d["N"])
        self.obstacles = np.ascontiguousarray(d["obstacles"], dtype=np.float64)
        self.n_obs = self.obstacles.shape[0]
        self.model = int(d.get("model", 0))     # 0: the scripts' 8-state UAV + camera; 1: the gimbal-less 5-state tracker (nmpc_b200.h)
        self.n_p = 8 if self.model else NP
        self.n_w, self.n_g = (3 if self.model else NU) * self.N, ((2 if self.model else 5) + self.n_obs) * (self.N + 1)
        self.device = device
        self._d = d
        self._h = C.c_void_p()
        self._max_batch = 0
        self._dev_cache: Dict[str, Any] = {}
        self._stats: Dict[str, Any] = {}
        self._warm = None
        self._create(max_batch)

    # -- handle management ---------------------------------------------------------------------
    def _create(self, max_batch: int):
        L = _ffi.lib()
        if self._h:
            L.nmpc_destroy(self._h)
            self._h = C.c_void_p()
        d = self._d
        spec = _ffi.NmpcSpec(self.T, self.N, self.n_obs, float(d.get("w1", 1.0)), float(d.get("w2", 2.0)),
                             float(d.get("vfov", 1.0)), float(d.get("hfov", 1.0)),
                             d["max_iter"], d["scaling"], d["tol"], int(max_batch), self.fill, self.model)
        _ffi.check(L.nmpc_create(C.byref(spec), self.device, C.byref(self._h)), "nmpc_create")
        self._max_batch = max_batch
        self.spec = spec

    def close(self):
        if self._h:
            _ffi.lib().nmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------
    def _dev_const(self, key: str, arr) -> "torch.Tensor":
        """Device copy of a batch-shared vector (bounds, obstacle table), cached by content."""
        if _is_cuda_tensor(arr):
            return arr.to(torch.float64).contiguous().view(-1)
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).reshape(-1))
        hit = self._dev_cache.get(key)
        if hit is not None and hit[0].shape == a.shape and np.array_equal(hit[0], a):
            return hit[1]
        t = torch.from_numpy(a.copy()).to(f"cuda:{self.device}")
        self._dev_cache[key] = (a.copy(), t)
        return t

    @staticmethod
    def _host(a, n: int, what: str = "argument", shared: bool = False) -> np.ndarray:
        """Host container -> contiguous float64 [rows, n].  Raises ValueError on a wrong length (the C ABI takes plain
        pointers: a short vector would be read past its end).  shared: a batch-shared vector (bounds), exactly n entries."""
        if torch is not None and isinstance(a, torch.Tensor):
            a = a.detach().cpu().numpy()
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 2 and a.shape[1] == 1 and a.shape[0] == n:   # CasADi column vector
            a = a.reshape(1, n)
        if a.ndim == 2 and a.shape[1] != n:
            raise ValueError(f"solver: {what} has {a.shape[1]} columns, expected {n}")
        if a.ndim != 2:
            if a.size % n != 0 or a.size == 0:
                raise ValueError(f"solver: {what} has {a.size} entries, expected a multiple of {n}")
            a = a.reshape(-1, n)
        if shared and a.shape[0] != 1:
            raise ValueError(f"solver: {what} must have exactly {n} entries (it is shared by the batch), got {a.size}")
        return np.ascontiguousarray(a)

    # -- the call --------------------------------------------------------------------------------
    def __call__(self, x0=None, p=None, lbx=None, ubx=None, lbg=None, ubg=None, obstacles=None,
                 want_g: bool = True, want_lam: bool = True, order=None, weights=None, target_traj=None,
                 blocking: bool = True, lam_x0=None, lam_g0=None):
        """lam_x0 [B, n_w], lam_g0 [B, n_g]: CasADi's multiplier guesses.  Like IPOPT they are ignored unless the solver was
        built with opts = {'ipopt': {'warm_start_init_point': 'yes', ...}} (no reference script does: non-reference fast mode,
        nmpc_set_warm_start); a NaN in lam_x0[b, 0] cold-starts instance b.
        weights: optional [B, 2] per-instance cost weights (w1, w2) for this call (numpy or CUDA tensor); the
        reference edits them in source (NMPC_TT.py:204-205) and its MATLAB outer loop sweeps them (MPC.m:90).
        target_traj: optional [B, N, 2] predicted target positions per stage (default: p[8:10] for every stage, as
        in the reference).
        blocking=False (host buffers only): enqueue the copies and the solve and return at once; the outputs are
        page-locked arrays owned by this solver that become valid after `solver.wait()` and are reused by the next
        call.  Pass page-locked inputs (torch pin_memory) for the copies to overlap."""
        if p is None:
            raise ValueError("solver: p is required")
        if any(v is None for v in (lbx, ubx, lbg, ubg)):
            raise ValueError("solver: lbx, ubx, lbg, ubg are required (the reference passes all four)")
        if self._d["warm"] and lam_x0 is not None and lam_g0 is not None:
            if not _is_cuda_tensor(p):       # the multiplier guesses live on the device: take the device path, hand numpy back
                dev = f"cuda:{self.device}"
                t = lambda a, n: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev).reshape(-1, n)
                B = self._batch_of(p)
                out = self(x0=t(x0, self.n_w), p=t(p, self.n_p), lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, obstacles=obstacles, want_g=want_g,
                           want_lam=want_lam, order=order, weights=weights, target_traj=target_traj, lam_x0=lam_x0, lam_g0=lam_g0)
                torch.cuda.synchronize()
                self._stats = {k: v.cpu().numpy() for k, v in self._stats.items()}
                return {k: (None if v is None else v.cpu().numpy()) for k, v in out.items()}
            B = self._batch_of(p)
            self.set_warm_start(torch.as_tensor(lam_x0, dtype=torch.float64, device=p.device).reshape(B, self.n_w),
                                torch.as_tensor(lam_g0, dtype=torch.float64, device=p.device).reshape(B, self.n_g))
            try:
                return self(x0=x0, p=p, lbx=lbx, ubx=ubx, lbg=lbg, ubg=ubg, obstacles=obstacles, want_g=want_g, want_lam=want_lam,
                            order=order, weights=weights, target_traj=target_traj)
            finally:
                self.set_warm_start(None, None)
        with self._weights(weights, p), self._traj(target_traj, p):
            if _is_cuda_tensor(p):
                return self._call_device(x0, p, lbx, ubx, lbg, ubg, obstacles, want_g, want_lam, order)
            return self._call_host(x0, p, lbx, ubx, lbg, ubg, obstacles, want_g, want_lam, blocking)

    def wait(self):
        """Complete a blocking=False call (nmpc_synchronize)."""
        _ffi.check(_ffi.lib().nmpc_synchronize(self._h), "nmpc_synchronize")

    def done(self) -> bool:
        """True when no blocking=False call is in flight any more (nmpc_query; never blocks)."""
        busy = C.c_int32(0)
        _ffi.check(_ffi.lib().nmpc_query(self._h, C.byref(busy)), "nmpc_query")
        return busy.value == 0

    def _pinned(self, key: str, shape, dtype=np.float64):
        """Page-locked output buffer, cached per (name, shape)."""
        k = ("pin", key, tuple(shape), np.dtype(dtype).str)
        buf = self._dev_cache.get(k)
        if buf is None:
            t = torch.empty(tuple(shape), dtype=torch.float64 if dtype == np.float64 else torch.int32).pin_memory()
            buf = (t, t.numpy())
            self._dev_cache[k] = buf
        return buf[1]

    def _batch_of(self, p) -> int:
        """Batch size of a parameter argument in any of the accepted containers (list, numpy, torch; (11,), (11,1), (B,11))."""
        if _is_cuda_tensor(p):
            return int(p.numel() // self.n_p)
        return int(np.asarray(p, dtype=np.float64).size // self.n_p)

    def _traj(self, traj, p):
        """Context manager: nmpc_set_target_trajectory for the duration of one call."""
        import contextlib
        if traj is None:
            return contextlib.nullcontext()
        B = self._batch_of(p)
        t = torch.as_tensor(traj if _is_cuda_tensor(traj) else np.asarray(traj, dtype=np.float64),
                            dtype=torch.float64, device=f"cuda:{self.device}").reshape(-1, self.N, 2).contiguous()
        if t.shape[0] != B:
            raise ValueError("solver: target_traj must be [B, N, 2]")
        L = _ffi.lib()

        @contextlib.contextmanager
        def cm():
            if B > self._max_batch:
                self._create(max(B, 2 * self._max_batch))
            _ffi.check(L.nmpc_set_target_trajectory(self._h, t.data_ptr()), "nmpc_set_target_trajectory")
            try:
                yield
            finally:
                self._keep_t = t            # the solve may still be in flight (device call, or blocking=False): hold the tensor
                L.nmpc_set_target_trajectory(self._h, None)
        return cm()

    def _weights(self, weights, p):
        """Context manager: nmpc_set_weights for the duration of one call."""
        import contextlib
        if weights is None:
            return contextlib.nullcontext()
        B = self._batch_of(p)
        w = torch.as_tensor(weights if _is_cuda_tensor(weights) else np.asarray(weights, dtype=np.float64),
                            dtype=torch.float64, device=f"cuda:{self.device}").reshape(-1, 2).contiguous()
        if w.shape[0] != B:
            raise ValueError("solver: weights must be [B, 2]")
        L = _ffi.lib()

        @contextlib.contextmanager
        def cm():
            if B > self._max_batch:          # grow first: the weights belong to the handle that runs the call
                self._create(max(B, 2 * self._max_batch))
            _ffi.check(L.nmpc_set_weights(self._h, w.data_ptr()), "nmpc_set_weights")
            try:
                yield
            finally:
                self._keep_w = w            # asynchronous call (device tensors, or blocking=False): the kernel reads w later
                L.nmpc_set_weights(self._h, None)
        return cm()

    def _obst(self, obstacles, B):
        if obstacles is None:
            return self.obstacles, 0
        o = obstacles.detach().cpu().numpy() if (torch is not None and isinstance(obstacles, torch.Tensor)) else obstacles
        o = np.ascontiguousarray(np.asarray(o, dtype=np.float64))
        if o.size == 3 * self.n_obs:
            return o.reshape(self.n_obs, 3), 0
        if o.size == 3 * self.n_obs * B:
            return o.reshape(B, self.n_obs, 3), _ffi.NMPC_OBS_PER_INSTANCE
        raise ValueError("solver: obstacles must be [n_obs,3] or [B,n_obs,3] of (cx, cy, r_uav + r_obs)")

    def _call_host(self, x0, p, lbx, ubx, lbg, ubg, obstacles, want_g, want_lam, blocking=True):
        L = _ffi.lib()
        single = np.asarray(p).ndim == 1 or (np.asarray(p).ndim == 2 and np.asarray(p).shape[1] == 1)
        p = self._host(p, self.n_p, "p")
        B = p.shape[0]
        x0 = np.zeros((B, self.n_w)) if x0 is None else self._host(x0, self.n_w, "x0")
        if x0.shape[0] != B:
            raise ValueError("solver: x0 and p disagree on the batch size")
        lbx, ubx = self._host(lbx, self.n_w, "lbx", True), self._host(ubx, self.n_w, "ubx", True)
        lbg, ubg = self._host(lbg, self.n_g, "lbg", True), self._host(ubg, self.n_g, "ubg", True)
        obs, flags = self._obst(obstacles, B)
        if B > self._max_batch:
            self._create(max(B, 2 * self._max_batch))
        new = (lambda k, shape, dt=np.float64: np.empty(shape, dtype=dt)) if blocking else self._pinned
        x = new("x", (B, self.n_w)); f = new("f", (B,))
        g = new("g", (B, self.n_g)) if want_g else None
        lam_x = new("lam_x", (B, self.n_w)) if want_lam else None
        lam_g = new("lam_g", (B, self.n_g)) if want_lam else None
        status = new("status", (B,), np.int32); iters = new("iters", (B,), np.int32)
        ptr = lambda a: None if a is None else a.ctypes.data
        fn = L.nmpc_solve_host if blocking else L.nmpc_solve_host_async
        if not blocking:
            self._keep = (p, x0, lbx, ubx, lbg, ubg, obs)      # inputs stay alive until wait()
        _ffi.check(fn(self._h, B, ptr(p), ptr(x0), ptr(lbx), ptr(ubx), ptr(lbg), ptr(ubg), ptr(obs), flags,
                      ptr(x), ptr(f), ptr(g), ptr(lam_x), ptr(lam_g), ptr(status), ptr(iters)),
                   "nmpc_solve_host")
        self._stats = dict(return_status=status, iter_count=iters)      # success is derived in stats(): valid after wait()
        def lam_p_host():      # CasADi's result dict also carries lam_p (nobody in the reference reads it): on first access
            dev = f"cuda:{self.device}"
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            xd, pd, ld, od = up(x), up(p), up(lam_g), up(obs)
            lp = torch.empty((B, self.n_p), dtype=torch.float64, device=dev)
            _ffi.check(L.nmpc_lam_p(self._h, B, xd.data_ptr(), pd.data_ptr(), od.data_ptr() if od.numel() else None, flags, ld.data_ptr(),
                                    lp.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "nmpc_lam_p")
            r = lp.cpu().numpy()
            return r[0] if single else r
        out = dict(x=x, f=f, g=g, lam_x=lam_x, lam_g=lam_g)
        if single:
            out = {k: (v[0] if v is not None else None) for k, v in out.items()}
        return _Sol(out, lam_p_host if (want_lam and blocking) else None)

    def _call_device(self, x0, p, lbx, ubx, lbg, ubg, obstacles, want_g, want_lam, order=None):
        L = _ffi.lib()
        dev = p.device
        single = p.dim() == 1
        p = p.to(torch.float64).reshape(-1, self.n_p).contiguous()
        B = p.shape[0]
        x0 = torch.zeros((B, self.n_w), dtype=torch.float64, device=dev) if x0 is None else x0.to(torch.float64).reshape(B, self.n_w).contiguous()
        lbx, ubx = self._dev_const("lbx", lbx), self._dev_const("ubx", ubx)
        lbg, ubg = self._dev_const("lbg", lbg), self._dev_const("ubg", ubg)
        flags = 0
        if obstacles is None:
            obs = self._dev_const("obs", self.obstacles)
        elif _is_cuda_tensor(obstacles):
            obs = obstacles.to(torch.float64).contiguous()
            flags = _ffi.NMPC_OBS_PER_INSTANCE if obs.numel() == 3 * self.n_obs * B and B > 1 else 0
        else:
            o, flags = self._obst(obstacles, B)
            obs = torch.from_numpy(o).to(dev)
        for t, n in ((lbx, self.n_w), (ubx, self.n_w), (lbg, self.n_g), (ubg, self.n_g)):
            if t.numel() != n:
                raise ValueError("solver: bound vector has the wrong length")
        x = torch.empty((B, self.n_w), dtype=torch.float64, device=dev)
        f = torch.empty(B, dtype=torch.float64, device=dev)
        g = torch.empty((B, self.n_g), dtype=torch.float64, device=dev) if want_g else None
        lam_x = torch.empty((B, self.n_w), dtype=torch.float64, device=dev) if want_lam else None
        lam_g = torch.empty((B, self.n_g), dtype=torch.float64, device=dev) if want_lam else None
        status = torch.empty(B, dtype=torch.int32, device=dev)
        iters = torch.empty(B, dtype=torch.int32, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        if order is not None:      # scheduling hint only: longest-first fetch order of the persistent kernel
            order = order.to(torch.int32).contiguous()
            if order.numel() != B:
                raise ValueError("solver: order must be a permutation of range(B)")
            _ffi.check(L.nmpc_set_order(self._h, order.data_ptr()), "nmpc_set_order")
        _ffi.check(L.nmpc_solve(self._h, B, ptr(p), ptr(x0), ptr(lbx), ptr(ubx), ptr(lbg), ptr(ubg), ptr(obs), flags,
                                ptr(x), ptr(f), ptr(g), ptr(lam_x), ptr(lam_g), ptr(status), ptr(iters), stream),
                   "nmpc_solve")
        def lam_p_dev():      # CasADi's result dict also carries lam_p (nobody in the reference reads it): on first access
            lp = torch.empty((B, self.n_p), dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                _ffi.check(L.nmpc_lam_p(self._h, B, ptr(x), ptr(p), ptr(obs), flags, ptr(lam_g), ptr(lp),
                                        torch.cuda.current_stream(dev).cuda_stream), "nmpc_lam_p")
            return lp[0] if single else lp
        self._keep = (p, x0, obs, order)   # keep inputs alive until the stream has consumed them
        self._stats = dict(return_status=status, iter_count=iters)      # success is derived lazily in stats()
        out = dict(x=x, f=f, g=g, lam_x=lam_x, lam_g=lam_g)
        if single:
            out = {k: (v[0] if v is not None else None) for k, v in out.items()}
        return _Sol(out, lam_p_dev if want_lam else None)

    def stats(self) -> Dict[str, Any]:
        """Per-instance outcome of the last call (CasADi: solver.stats()); after a blocking=False call, valid once
        wait() has returned."""
        d = dict(self._stats)
        if d and "success" not in d:
            d["success"] = d["return_status"] == 0
        return d

    def work_counters(self) -> Dict[str, int]:
        st = _ffi.NmpcStats()
        _ffi.check(_ffi.lib().nmpc_get_stats(self._h, C.byref(st)), "nmpc_get_stats")
        return {name: int(getattr(st, name)) for name, _ in _ffi.NmpcStats._fields_}

    # -- function-level evaluation (nlp_f / nlp_g / nlp_grad_f / nlp_hess_l of the reference's nlpsol) ----
    def evaluate(self, w, p, lam=None, v=None, sigma: float = 1.0, obstacles=None, weights=None, target_traj=None):
        if weights is not None or target_traj is not None:
            pp = np.asarray(p) if not _is_cuda_tensor(p) else p
            with self._weights(weights, pp), self._traj(target_traj, pp):
                return self.evaluate(w, p, lam=lam, v=v, sigma=sigma, obstacles=obstacles)
        L = _ffi.lib()
        dev = f"cuda:{self.device}"
        to = lambda a, n: torch.as_tensor(np.asarray(a, dtype=np.float64) if not _is_cuda_tensor(a) else a,
                                          dtype=torch.float64, device=dev).reshape(-1, n).contiguous()
        w, p = to(w, self.n_w), to(p, self.n_p)
        B = w.shape[0]
        lam_t = to(lam, self.n_g) if lam is not None else None
        v_t = to(v, self.n_w) if v is not None else None
        flags = 0
        if obstacles is None:
            obs = self._dev_const("obs", self.obstacles)
        else:
            o, flags = self._obst(obstacles, B)
            obs = torch.from_numpy(o).to(dev)
        z = lambda *s: torch.empty(s, dtype=torch.float64, device=dev)
        f, g, grad = z(B), z(B, self.n_g), z(B, self.n_w)
        jtv = z(B, self.n_w) if lam_t is not None else None
        hv = z(B, self.n_w) if v_t is not None else None
        ptr = lambda t: None if t is None else t.data_ptr()
        stream = torch.cuda.current_stream(torch.device(dev)).cuda_stream
        _ffi.check(L.nmpc_eval(self._h, B, ptr(w), ptr(p), ptr(obs), flags, float(sigma), ptr(lam_t), ptr(v_t),
                               ptr(f), ptr(g), ptr(grad), ptr(jtv), ptr(hv), stream), "nmpc_eval")
        torch.cuda.synchronize()
        return dict(f=f, g=g, grad=grad, jtv=jtv, hv=hv)

    # -- shift_timestep (NMPC_TT.py:13-30) on device --------------------------------------------
    def solve_and_step(self, p, u_warm, lbx, ubx, lbg, ubg, target_vw, fov_centre=None, err_accum=None, obstacles=None,
                       want_x: bool = False, weights=None, target_traj=None):
        if weights is not None or target_traj is not None:
            with self._weights(weights, p), self._traj(target_traj, p):
                return self.solve_and_step(p, u_warm, lbx, ubx, lbg, ubg, target_vw, fov_centre, err_accum, obstacles, want_x)
        return self._solve_and_step(p, u_warm, lbx, ubx, lbg, ubg, target_vw, fov_centre, err_accum, obstacles, want_x)

    def _solve_and_step(self, p, u_warm, lbx, ubx, lbg, ubg, target_vw, fov_centre, err_accum, obstacles, want_x):
        """One closed-loop step of B instances in ONE kernel launch (nmpc_solve_and_step): solves from the warm start
        u_warm with parameters p, then shifts p / u_warm in place (torch CUDA float64 tensors).  Returns dict(x, f)
        with x None unless want_x; status / iterations through stats()."""
        L = _ffi.lib()
        dev = p.device
        B = p.shape[0]
        lbx, ubx = self._dev_const("lbx", lbx), self._dev_const("ubx", ubx)
        lbg, ubg = self._dev_const("lbg", lbg), self._dev_const("ubg", ubg)
        flags = 0
        if obstacles is None:
            obs = self._dev_const("obs", self.obstacles)
        else:
            obs = obstacles.to(torch.float64).contiguous()
            flags = _ffi.NMPC_OBS_PER_INSTANCE if obs.numel() == 3 * self.n_obs * B and B > 1 else 0
        x = torch.empty((B, self.n_w), dtype=torch.float64, device=dev) if want_x else None
        f = torch.empty(B, dtype=torch.float64, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        iters = torch.empty(B, dtype=torch.int32, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        if target_vw is None and getattr(self, "_sched", None) is None:
            raise ValueError("solver: target_vw is None and no schedule is set (set_schedule)")
        _ffi.check(L.nmpc_solve_and_step(self._h, B, ptr(p), ptr(u_warm), ptr(lbx), ptr(ubx), ptr(lbg), ptr(ubg), ptr(obs), flags,
                                         ptr(target_vw), ptr(x), ptr(f), ptr(fov_centre), ptr(err_accum), ptr(status), ptr(iters),
                                         stream), "nmpc_solve_and_step")
        self._keep = (obs,)
        self._stats = dict(return_status=status, iter_count=iters)
        return dict(x=x, f=f, g=None, lam_x=None, lam_g=None)

    def run_closed_loop(self, steps: int, p, u_warm, lbx, ubx, lbg, ubg, target_vw, fov_centre=None, err_accum=None,
                        obstacles=None, want_x: bool = False, log: bool = True):
        """`steps` closed-loop steps of every instance in ONE launch, without a batch-wide barrier between the steps
        (nmpc_run_closed_loop; the scripts' `while mpc_iter < sim_time / T` loop, NMPC_TT.py:346-402).  p / u_warm are advanced in
        place.  Returns dict(x, f of the last step, status_log / iters_log [steps, B] when log, converged [B] = number of
        converged solves per instance); stats() reports the last step.  Bit-identical to `steps` solve_and_step calls."""
        L = _ffi.lib()
        dev = p.device
        B = p.shape[0]
        lbx, ubx = self._dev_const("lbx", lbx), self._dev_const("ubx", ubx)
        lbg, ubg = self._dev_const("lbg", lbg), self._dev_const("ubg", ubg)
        flags = 0
        if obstacles is None:
            obs = self._dev_const("obs", self.obstacles)
        else:
            obs = obstacles.to(torch.float64).contiguous()
            flags = _ffi.NMPC_OBS_PER_INSTANCE if obs.numel() == 3 * self.n_obs * B and B > 1 else 0
        if target_vw is None and getattr(self, "_sched", None) is None:
            raise ValueError("solver: target_vw is None and no schedule is set (set_schedule)")
        x = torch.empty((B, self.n_w), dtype=torch.float64, device=dev) if want_x else None
        f = torch.empty(B, dtype=torch.float64, device=dev)
        slog = torch.empty((steps, B), dtype=torch.int32, device=dev) if log else None
        ilog = torch.empty((steps, B), dtype=torch.int32, device=dev) if log else None
        conv = torch.zeros(B, dtype=torch.int32, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _ffi.check(L.nmpc_run_closed_loop(self._h, B, int(steps), ptr(p), ptr(u_warm), ptr(lbx), ptr(ubx), ptr(lbg), ptr(ubg), ptr(obs), flags,
                                          ptr(target_vw), ptr(x), ptr(f), ptr(fov_centre), ptr(err_accum), ptr(slog), ptr(ilog), ptr(conv),
                                          stream), "nmpc_run_closed_loop")
        self._keep = (obs,)
        self._stats = dict(return_status=slog[-1], iter_count=ilog[-1]) if log else {}
        return dict(x=x, f=f, status_log=slog, iters_log=ilog, converged=conv)

    def set_warm_start(self, lam_x0, lam_g0):
        """nmpc_set_warm_start: CUDA float64 tensors lam_x0 [B, n_w], lam_g0 [B, n_g] read by the following solves (and
        overwritten with the shifted multipliers by solve_and_step); None, None restores IPOPT's cold multiplier start.
        NON-REFERENCE mode; the pushes / mu_init come from the solver's 'ipopt' options."""
        L = _ffi.lib()
        if lam_x0 is None:
            self._warm = None
            _ffi.check(L.nmpc_set_warm_start(self._h, None, None, None), "nmpc_set_warm_start")
            return
        if not (_is_cuda_tensor(lam_x0) and _is_cuda_tensor(lam_g0)) or lam_x0.dtype != torch.float64 or lam_g0.dtype != torch.float64:
            raise ValueError("solver: lam_x0 / lam_g0 must be CUDA float64 tensors")
        lam_x0, lam_g0 = lam_x0.contiguous(), lam_g0.contiguous()
        if lam_x0.numel() % self.n_w or lam_g0.numel() % self.n_g or lam_x0.numel() // self.n_w != lam_g0.numel() // self.n_g:
            raise ValueError("solver: lam_x0 must be [B, n_w] and lam_g0 [B, n_g]")
        self._warm = (lam_x0, lam_g0)        # owned here: the handle keeps raw pointers
        o = _ffi.NmpcWarmOpts(*self._d["warm_opts"])
        _ffi.check(L.nmpc_set_warm_start(self._h, lam_x0.data_ptr(), lam_g0.data_ptr(), C.byref(o)), "nmpc_set_warm_start")

    def set_schedule(self, table, row_of_instance=None, phase=None, mpc_iter: int = 0):
        """Device-side target schedule (nmpc_set_schedule): table [n_rows, len, 2] of (v, omega) per closed-loop step;
        instance b follows row row_of_instance[b] starting at step phase[b].  solve_and_step(target_vw=None) then looks
        the pair up in the kernel's epilogue and advances the step counter.  table=None removes the schedule."""
        L = _ffi.lib()
        if table is None:
            self._sched = None
            _ffi.check(L.nmpc_set_schedule(self._h, None, 0, 0, None, None, 0), "nmpc_set_schedule")
            return
        dev = f"cuda:{self.device}"
        t = torch.as_tensor(np.asarray(table, dtype=np.float64) if not torch.is_tensor(table) else table, dtype=torch.float64, device=dev).contiguous()
        if t.dim() != 3 or t.shape[2] != 2:
            raise ValueError("solver: schedule table must be [n_rows, len, 2]")
        i32 = lambda a: None if a is None else torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a, device=dev).to(torch.int32).contiguous()
        rows, ph = i32(row_of_instance), i32(phase)
        if rows is not None and (int(rows.min()) < 0 or int(rows.max()) >= t.shape[0]):
            raise ValueError("solver: schedule row index out of range")
        if ph is not None and int(ph.min()) < 0:
            raise ValueError("solver: schedule phases must be >= 0")
        self._sched = (t, rows, ph)          # owned here: the handle keeps raw pointers
        ptr = lambda a: None if a is None else a.data_ptr()
        _ffi.check(L.nmpc_set_schedule(self._h, t.data_ptr(), int(t.shape[0]), int(t.shape[1]), ptr(rows), ptr(ph), int(mpc_iter)),
                   "nmpc_set_schedule")

    def step(self, x_sol, p, u_warm, target_vw, fov_centre=None, err_accum=None):
        """In-place closed-loop shift of B instances (torch CUDA float64 tensors): p [B,11], u_warm [B,6N];
        err_accum [B] += ||new FOV centre - this step's target|| (NMPC_TT.py:435)."""
        L = _ffi.lib()
        B = p.shape[0]
        stream = torch.cuda.current_stream(p.device).cuda_stream
        _ffi.check(L.nmpc_step(self._h, B, x_sol.data_ptr(), p.data_ptr(), u_warm.data_ptr(),
                               target_vw.data_ptr(), None if fov_centre is None else fov_centre.data_ptr(),
                               None if err_accum is None else err_accum.data_ptr(), stream),
                   "nmpc_step")


def nlpsol(name: str, plugin: str, problem, opts: Optional[dict] = None, device: int = 0, max_batch: int = 1,
           fill: int = 1) -> Solver:
    """Mirror of ca.nlpsol(name, 'ipopt', nlp_prob, opts) (NMPC_TT.py:267).

    plugin: 'ipm' (or 'ipopt' for drop-in spelling) -- the in-kernel interior-point method.
    problem: a scenarios.Scenario or a dict(T=, N=, obstacles=[(cx,cy,r_obs)...], uav_r=5, w1=1, w2=2).
    opts: {'ipopt': {'max_iter': 100, ...}} as in the reference; unknown keys are ignored like print_level.
    fill: scheduling hint for solvers that run next to others (PipelinedClosedLoop): instances per resident warp."""
    if plugin not in ("ipm", "ipopt"):
        raise ValueError(f"nlpsol: unknown plugin '{plugin}' (only the in-kernel 'ipm' exists)")
    return Solver(name, problem, opts, device=device, max_batch=max_batch, fill=fill)
