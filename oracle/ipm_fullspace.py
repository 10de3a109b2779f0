"""TEST INFRASTRUCTURE ONLY (oracle) -- a second, structurally independent restatement of IPOPT's main loop.

PARITY UNPINNED (as oracle/nmpc_oracle.cpp): no CasADi/IPOPT here, so this file, too, restates the published algorithm
(Waechter & Biegler, Math. Prog. 106 (2006), with the option defaults of IPOPT 3.12 and the options the scripts set,
Python/NMPC_TT.py:257-265).  Its purpose is to catch what comparing the C++ oracle with the CUDA kernel cannot: those two
share the condensed linear algebra (eliminated slacks and multipliers, "inertia correct <=> reduced matrix positive
definite") and hand-written or jet derivatives.  Here instead

  * the primal-dual step comes from IPOPT's FULL-SPACE augmented system
        [ W + Sigma_x + dw I        0            J^T ] [dx]     [ grad phi_x + J^T y ]
        [        0           Sigma_s + dw I      -I  ] [ds] = - [ grad phi_s - y     ]
        [        J                 -I             0  ] [dy]     [ d(x) - s           ]
    (346 x 346 for the three-obstacle scripts), solved by a dense LU, and the inertia IPOPT asks MUMPS for is COUNTED
    from the pivots of a Bunch-Kaufman L D L^T factorisation of that symmetric indefinite matrix: (n + m, m, 0), or the
    perturbation dw is raised by IPOPT's rule;
  * f, g, grad f, J and the Lagrangian Hessian are torch.autograd derivatives of the literal formulas in
    oracle/nlp_ref.py (the role CasADi's AD plays in the reference);
  * the algorithm state is a handful of numpy vectors; nothing is shared with nmpc_oracle.cpp.

Restated: gradient-based scaling, bound relaxation, DefaultIterateInitializer with least-squares multipliers, monotone
barrier update, fraction to the boundary, inertia correction, the filter line search with second-order corrections and
the filter-reset heuristic, the watchdog, the soft restoration phase, tiny-step handling, the kappa_sigma multiplier
safeguard, the scaled-error convergence test.  NOT restated: the restoration phase proper (a solve that needs it stops
with status "needs_restoration" -- the test checks that this happens at the iteration where the C++ oracle enters it) and
the slack safeguard ("needs_slack_safeguard").  tests/test_oracle_solve.py::test_fullspace_ipm_reproduces_oracle_iterates
compares iteration logs with the C++ oracle.  Only tests/ may import this module.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg
import torch

from . import nlp_ref

EPS = np.finfo(np.float64).eps


class Problem:
    """The reference NLP of one instance: literal formulas + autograd (nlp_ref), numpy in / numpy out."""

    def __init__(self, spec, p):
        """spec: nlp_ref.RefSpec (the Python scripts' NLP) or nlp_ref.RefSpec5 (MATLAB/Dynamic Obstacles/NMPC_TT.m)."""
        self.spec = spec
        self.p = torch.tensor(np.asarray(p, dtype=np.float64))
        self.n, self.m = spec.n_w, spec.n_g
        five = isinstance(spec, nlp_ref.RefSpec5)
        self._f = nlp_ref.objective5 if five else nlp_ref.objective
        self._g = nlp_ref.constraints5 if five else nlp_ref.constraints

    def fg(self, x):
        with torch.no_grad():
            w = torch.tensor(x)
            return float(self._f(self.spec, w, self.p)), self._g(self.spec, w, self.p).numpy()

    def derivs(self, x):
        w = torch.tensor(x, requires_grad=True)
        grad = torch.autograd.grad(self._f(self.spec, w, self.p), w)[0].numpy()
        J = torch.func.jacrev(lambda ww: self._g(self.spec, ww, self.p))(torch.tensor(x)).numpy()
        return grad, J

    def hess(self, x, sigma, lam):
        lt = torch.tensor(lam)
        H = torch.autograd.functional.hessian(
            lambda ww: sigma * self._f(self.spec, ww, self.p) + (lt * self._g(self.spec, ww, self.p)).sum(),
            torch.tensor(x), vectorize=True)
        return H.numpy()


def _cmp_le(lhs, rhs, base):
    # IPOPT's Compare_le: lhs <= rhs up to ten machine epsilons of the base value
    return lhs - rhs <= 10.0 * EPS * abs(base)


def _inertia(K):
    """(positive, negative, zero) eigenvalue counts of the symmetric matrix K from a Bunch-Kaufman L D L^T factorisation
    (Sylvester's law of inertia: the counts of K are those of the block-diagonal D) -- what IPOPT obtains from MUMPS."""
    _, D, _ = scipy.linalg.ldl(K, lower=True)
    pos = neg = zero = 0
    k, N = 0, K.shape[0]
    while k < N:
        if k + 1 < N and D[k + 1, k] != 0.0:                     # 2 x 2 pivot
            ev = np.linalg.eigvalsh(D[k:k + 2, k:k + 2]); k += 2
        else:
            ev = D[k:k + 1, k]; k += 1
        pos += int((ev > 0).sum()); neg += int((ev < 0).sum()); zero += int((ev == 0).sum())
    return pos, neg, zero


def solve(prob: Problem, x0, lbx, ubx, lbg, ubg, *, max_iter=100, tol=1e-8, log=None, events=None):
    """One cold-started solve.  Returns dict(status, iters, x, f, lam_x, lam_g); log (a list) receives one tuple
    (mu, f, inf_pr, inf_du, dw, alpha_pr, alpha_du, trial points) per iteration, as IPOPT's iteration output would;
    events (a list) receives (iteration, what) whenever a globalisation heuristic acts."""
    n, m = prob.n, prob.m
    # options (IPOPT 3.12 defaults unless the scripts set them)
    dual_inf_tol, constr_viol_tol, compl_inf_tol = 1.0, 1e-4, 1e-4
    bound_relax, bound_push, bound_frac = 1e-8, 1e-2, 1e-2
    mu, kappa_mu, theta_mu, kappa_eps, tau_min = 0.1, 0.2, 1.5, 10.0, 0.99
    kappa_d, kappa_sigma, s_max = 1e-4, 1e10, 100.0
    max_grad, scal_min, y_init_max = 100.0, 1e-8, 1e3
    dw_first, dw_min, dw_max, k_first, k_inc, k_dec = 1e-4, 1e-20, 1e20, 100.0, 8.0, 1.0 / 3.0
    g_theta, g_phi, eta_phi, s_theta, s_phi, delta = 1e-5, 1e-8, 1e-8, 1.1, 2.3, 1.0
    a_min_frac, max_soc, kappa_soc, obj_max_inc = 0.05, 4, 0.99, 5.0

    x = np.array(x0, dtype=np.float64)
    lbx, ubx, lbg, ubg = (np.asarray(a, dtype=np.float64) for a in (lbx, ubx, lbg, ubg))

    # ---- NLP scaling from the derivatives at the user's starting point (GradientScaling)
    grad_u, J_u = prob.derivs(x)
    gmax = np.abs(grad_u).max()
    df = max(scal_min, max_grad / gmax) if gmax > max_grad else 1.0
    rmax = np.abs(J_u).max(axis=1)
    dc = np.where(rmax > max_grad, np.maximum(scal_min, max_grad / np.maximum(rmax, 1e-300)), 1.0)

    def relax(b, sign):
        fin = np.abs(b) < 1e19
        return np.where(fin, b + sign * bound_relax * np.maximum(1.0, np.abs(b)), sign * np.inf)
    xL, xU = relax(lbx, -1.0), relax(ubx, 1.0)
    sL = relax(np.where(np.abs(lbg) < 1e19, dc * lbg, lbg), -1.0)
    sU = relax(np.where(np.abs(ubg) < 1e19, dc * ubg, ubg), 1.0)
    LO = np.concatenate([np.isfinite(xL), np.isfinite(sL)])      # which of the n + m primal variables (x, s) has a lower bound
    UP = np.concatenate([np.isfinite(xU), np.isfinite(sU)])
    lo = np.concatenate([xL, sL]); up = np.concatenate([xU, sU])
    one_lo, one_up = LO & ~UP, UP & ~LO

    def fg(xv):
        f_, g_ = prob.fg(xv)
        return df * f_, dc * g_

    def derivs(xv):
        gr, J_ = prob.derivs(xv)
        return df * gr, dc[:, None] * J_

    # ---- starting point (DefaultIterateInitializer)
    def push(v, l, u):
        v = v.copy()
        both = np.isfinite(l) & np.isfinite(u)
        l, u = np.where(np.isfinite(l), l, -1e300), np.where(np.isfinite(u), u, 1e300)     # keep the unused branches finite
        pl = np.minimum(bound_push * np.maximum(1.0, np.abs(l)), bound_frac * (u - l))
        pu = np.minimum(bound_push * np.maximum(1.0, np.abs(u)), bound_frac * (u - l))
        v[both] = np.minimum(np.maximum(v[both], (l + pl)[both]), (u - pu)[both])
        ol = (l > -1e300) & ~(u < 1e300)
        v[ol] = np.maximum(v[ol], (l + bound_push * np.maximum(1.0, np.abs(l)))[ol])
        ou = (u < 1e300) & ~(l > -1e300)
        v[ou] = np.minimum(v[ou], (u - bound_push * np.maximum(1.0, np.abs(u)))[ou])
        return v
    x = push(x, xL, xU)
    f, g = fg(x)
    s = push(g, sL, sU)
    zl = LO.astype(np.float64); zu = UP.astype(np.float64)       # bound multipliers of (x, s): z_L, v_L | z_U, v_U
    grad, J = derivs(x)
    A = np.hstack([J, -np.eye(m)])
    b = np.concatenate([grad, np.zeros(m)]) - zl + zu
    K = np.block([[np.eye(n + m), A.T], [A, np.zeros((m, m))]])
    y = np.linalg.solve(K, np.concatenate([-b, np.zeros(m)]))[n + m:]
    if not np.abs(y).max() <= y_init_max:
        y = np.zeros(m)

    def slacks(xv, sv):
        v = np.concatenate([xv, sv])
        with np.errstate(invalid="ignore"):
            return np.where(LO, v - lo, 1.0), np.where(UP, up - v, 1.0)

    def barrier(xv, sv, fv, mu_):
        a, c = slacks(xv, sv)
        if (a[LO] <= 0).any() or (c[UP] <= 0).any():
            return math.inf
        return fv - mu_ * (np.log(a[LO]).sum() + np.log(c[UP]).sum()) + kappa_d * mu_ * (a[one_lo].sum() + c[one_up].sum())

    def errors(xv, sv, yv, zl_, zu_, gv, gradv, Jv, mu_):
        a, c = slacks(xv, sv)
        du = max(np.abs(gradv + Jv.T @ yv - zl_[:n] + zu_[:n]).max(), np.abs(-yv - zl_[n:] + zu_[n:]).max())
        pr = np.abs(gv - sv).max()
        co = max(np.abs(a * zl_ - mu_)[LO].max(initial=0.0), np.abs(c * zu_ - mu_)[UP].max(initial=0.0))
        nb = int(LO.sum() + UP.sum())
        sumz = zl_[LO].sum() + zu_[UP].sum()
        sd = max(s_max, (np.abs(yv).sum() + sumz) / max(1, m + nb)) / s_max
        sc = max(s_max, sumz / max(1, nb)) / s_max
        return du, pr, co, max(du / sd, pr, co / sc)

    def converged():
        du, pr, co, E = errors(x, s, y, zl, zu, g, grad, J, 0.0)
        gu = g / dc
        viol = max(np.where(np.abs(lbg) < 1e19, lbg - gu, -np.inf).max(), np.where(np.abs(ubg) < 1e19, gu - ubg, -np.inf).max(), 0.0)
        return E <= tol and du / df <= dual_inf_tol and viol <= constr_viol_tol and co / df <= compl_inf_tol

    # ---------------------------------------------------------------------------------------------------------------
    # algorithm state.  P = the current iterate with everything evaluated at it.
    P = dict(x=x, s=s, y=y, zl=zl, zu=zu, f=f, g=g, grad=grad, J=J)
    filt: list[tuple[float, float]] = []
    theta_max = theta_min = None
    dw_last = 0.0
    tau = max(tau_min, 1.0 - mu)
    last_rej_filter, succ_filter_rej, n_filter_resets = False, 0, 0            # filter reset heuristic
    in_watchdog, wd_short, wd_trial, wd = False, 0, 0, None                      # watchdog
    in_soft, soft_cnt = False, 0                                                 # soft restoration phase
    tiny_last, tiny_flag = False, False                                          # tiny-step handling
    counters = dict(watchdog_starts=0, soft_resto_steps=0, filter_resets=0, tiny_steps=0)
    it = 0

    def ftb(vals, steps, mask, tau_):
        neg = mask & (steps < 0)
        return min(1.0, (-tau_ * vals[neg] / steps[neg]).min(initial=1.0))

    def err_at(Q, mu_):
        return errors(Q["x"], Q["s"], Q["y"], Q["zl"], Q["zu"], Q["g"], Q["grad"], Q["J"], mu_)

    def pd_error(Q, mu_):
        """IpoptCalculatedQuantities::*_primal_dual_system_error: averaged 1-norms (soft restoration test)"""
        a_, c_ = slacks(Q["x"], Q["s"])
        du = np.abs(Q["grad"] + Q["J"].T @ Q["y"] - Q["zl"][:n] + Q["zu"][:n]).sum() + np.abs(-Q["y"] - Q["zl"][n:] + Q["zu"][n:]).sum()
        pr = np.abs(Q["g"] - Q["s"]).sum()
        co = np.abs(a_ * Q["zl"] - mu_)[LO].sum() + np.abs(c_ * Q["zu"] - mu_)[UP].sum()
        nb = int(LO.sum() + UP.sum())
        return du / (n + m) + pr / m + co / nb

    while True:
        # ---- OptimalityErrorConvergenceCheck
        du0, pr0, co0, E0 = err_at(P, 0.0)
        gu = P["g"] / dc
        viol = max(np.where(np.abs(lbg) < 1e19, lbg - gu, -np.inf).max(), np.where(np.abs(ubg) < 1e19, gu - ubg, -np.inf).max(), 0.0)
        if E0 <= tol and du0 / df <= dual_inf_tol and viol <= constr_viol_tol and co0 / df <= compl_inf_tol:
            status = "Solve_Succeeded"; break
        if it >= max_iter:
            status = "Maximum_Iterations_Exceeded"; break
        # ---- MonotoneMuUpdate (mu_allow_fast_monotone_decrease); a second tiny step in a row forces mu down
        mu_floor = min(tol, df * compl_inf_tol) / (kappa_eps + 1.0)
        tf, tiny_flag, done = tiny_flag, False, False
        while (err_at(P, mu)[3] <= kappa_eps * mu or tf) and not done:
            new = max(mu_floor, min(kappa_mu * mu, mu ** theta_mu))
            if new == mu:
                if tf:
                    return dict(status="Search_Direction_Becomes_Too_Small", iters=it, counters=counters)
                break
            mu = new; tau = max(tau_min, 1.0 - mu)
            if tf:
                done, tf = True, False
            in_soft, soft_cnt = False, 0                                         # BacktrackingLineSearch::Reset
            filt, last_rej_filter, succ_filter_rej = [], False, 0                # FilterLSAcceptor::Reset
        # ---- search direction from the full-space system, inertia counted from the LDL^T pivots
        x, s, y, zl, zu, f, g, grad, J = (P[k] for k in ("x", "s", "y", "zl", "zu", "f", "g", "grad", "J"))
        a, c = slacks(x, s)
        sig = np.where(LO, zl / a, 0.0) + np.where(UP, zu / c, 0.0)
        gphi = np.concatenate([grad, np.zeros(m)]) - np.where(LO, mu / a, 0.0) + np.where(UP, mu / c, 0.0) \
            + kappa_d * mu * (one_lo.astype(np.float64) - one_up.astype(np.float64))
        W = prob.hess(x, df, dc * y)
        A = np.hstack([J, -np.eye(m)])
        rhs1 = gphi + A.T @ y
        dw = 0.0
        while True:
            K = np.zeros((n + 2 * m, n + 2 * m))
            K[:n, :n] = W
            K[np.arange(n + m), np.arange(n + m)] += sig + dw
            K[:n + m, n + m:] = A.T
            K[n + m:, :n + m] = A
            if _inertia(K) == (n + m, m, 0):
                break
            if dw == 0.0:
                dw = dw_first if dw_last == 0.0 else max(dw_min, dw_last * k_dec)
            else:
                dw = (k_first if (dw_last == 0.0 or 1e5 * dw_last < dw) else k_inc) * dw
            if dw > dw_max:
                return dict(status="needs_restoration", iters=it, why="inertia", counters=counters)
        if dw > 0:
            dw_last = dw

        def direction(cres_, K=K, rhs1=rhs1, a=a, c=c, zl=zl, zu=zu, mu=mu):
            sol = np.linalg.solve(K, -np.concatenate([rhs1, cres_]))
            dv_, dy_ = sol[:n + m], sol[n + m:]
            return dict(v=dv_, y=dy_, zl=np.where(LO, (mu - zl * dv_) / a - zl, 0.0), zu=np.where(UP, (mu + zu * dv_) / c - zu, 0.0))

        cres = g - s
        step = direction(cres)

        def a_max_of(stp, Q=None):
            Q = P if Q is None else Q                                            # (P is looked up at call time: the watchdog may restore it)
            a_, c_ = slacks(Q["x"], Q["s"])
            return min(ftb(a_, stp["v"], LO, tau), ftb(c_, -stp["v"], UP, tau))

        def a_dual_of(stp, Q=None):
            Q = P if Q is None else Q
            return min(ftb(Q["zl"], stp["zl"], LO, tau), ftb(Q["zu"], stp["zu"], UP, tau))

        # ---- BacktrackingLineSearch::FindAcceptableTrialPoint ---------------------------------------------------------
        cur_theta = np.abs(cres).sum()
        if not in_watchdog:                                                      # InitThisLineSearch
            if n_filter_resets < 5:                                              # filter reset heuristic (max_filter_resets = 5)
                if last_rej_filter:
                    succ_filter_rej += 1
                    if succ_filter_rej >= 5:                                     # filter_reset_trigger
                        filt, last_rej_filter, succ_filter_rej = [], False, 0
                        n_filter_resets += 1; counters["filter_resets"] += 1
                        if events is not None:
                            events.append((it, "filter_reset"))
                else:
                    succ_filter_rej = 0
            last_rej_filter = False
            ref = dict(theta=cur_theta, phi=barrier(x, s, f, mu), gbd=float(gphi @ step["v"]))
        else:
            ref = wd["ref"]
        if theta_max is None:
            theta_max = 1e4 * max(1.0, ref["theta"]); theta_min = 1e-4 * max(1.0, ref["theta"])

        def ftype(al):
            if ref["theta"] == 0.0 and 0.0 < ref["gbd"] < 100.0 * EPS:
                return True
            return ref["gbd"] < 0 and al * (-ref["gbd"]) ** s_phi > delta * ref["theta"] ** s_theta

        def armijo(al, phi_t):
            return _cmp_le(phi_t - ref["phi"], eta_phi * al * ref["gbd"], ref["phi"])

        def acceptable(al, phi_t, th_t):
            nonlocal last_rej_filter
            if not (math.isfinite(phi_t) and math.isfinite(th_t)) or th_t > theta_max:
                return False
            if al > 0.0 and ftype(al) and ref["theta"] <= theta_min:
                ok = armijo(al, phi_t)
            else:
                ok = True
                if phi_t > ref["phi"]:
                    bas = math.log10(abs(ref["phi"])) if abs(ref["phi"]) > 10.0 else 1.0
                    ok = not (math.log10(phi_t - ref["phi"]) > obj_max_inc + bas)
                ok = ok and (_cmp_le(th_t, (1 - g_theta) * ref["theta"], ref["theta"]) or _cmp_le(phi_t - ref["phi"], -g_phi * ref["theta"], ref["phi"]))
            if not ok:
                last_rej_filter = False
                return False
            ok = all(_cmp_le(phi_t, fb, fb) or _cmp_le(th_t, ft, ft) for fb, ft in filt)
            if not ok:
                last_rej_filter = True
            return ok

        def augment_filter():
            nonlocal filt
            ne = (ref["phi"] - g_phi * ref["theta"], (1 - g_theta) * ref["theta"])
            filt = [e for e in filt if not (e[0] >= ne[0] and e[1] >= ne[1])] + [ne]

        def trial(al, stp, Q=None):
            Q = P if Q is None else Q
            xt = Q["x"] + al * stp["v"][:n]; st_ = Q["s"] + al * stp["v"][n:]
            ft, gt = fg(xt)
            return dict(x=xt, s=st_, f=ft, g=gt, theta=np.abs(gt - st_).sum(), phi=barrier(xt, st_, ft, mu))

        n_trials = 0

        def backtrack(stp, Q, skip_first):
            """DoBacktrackingLineSearch from iterate Q along stp.  Returns (trial point or None, alpha_primal, step used, shortened)."""
            nonlocal n_trials
            a_max = a_max_of(stp, Q)
            if in_watchdog:
                a_min, a_test = a_max, wd["alpha_test"]
            else:
                a_min = g_theta
                if ref["gbd"] < 0:
                    a_min = min(g_theta, g_phi * ref["theta"] / (-ref["gbd"]))
                    if ref["theta"] <= theta_min:
                        a_min = min(a_min, delta * ref["theta"] ** s_theta / (-ref["gbd"]) ** s_phi)
                a_min *= a_min_frac
                a_test = a_max
            alpha, shortened, got, used = a_max, 0, None, stp
            if skip_first:
                alpha *= 0.5
            while alpha > a_min or shortened == 0:
                t = trial(alpha, stp, Q); n_trials += 1
                if not in_watchdog:
                    a_test = alpha
                if acceptable(a_test, t["phi"], t["theta"]):
                    got = t; break
                if in_watchdog:
                    break
                if alpha == a_max and ref["theta"] <= t["theta"] and math.isfinite(t["phi"]):
                    # second-order correction: same matrix, constraint residual alpha * c + c(trial), accumulated
                    csoc, a_soc, th_old, cnt, ts = cres.copy(), alpha, 0.0, 0, t
                    while cnt < max_soc and got is None and (cnt == 0 or ts["theta"] <= kappa_soc * th_old):
                        th_old = ts["theta"]
                        csoc = a_soc * csoc + (ts["g"] - ts["s"])
                        sstep = direction(csoc)
                        a_soc = a_max_of(sstep, Q)
                        ts = trial(a_soc, sstep, Q); n_trials += 1
                        if not math.isfinite(ts["phi"]):
                            break
                        if acceptable(a_test, ts["phi"], ts["theta"]):
                            got, used, alpha = ts, sstep, a_soc
                        else:
                            cnt += 1
                    if got is not None:
                        break
                alpha *= 0.5; shortened += 1
            if got is not None and not (ftype(a_test) and armijo(a_test, got["phi"])):
                augment_filter()                                                 # not an Armijo-accepted f-type step
            return got, alpha, used, shortened

        def soft_step(stp):
            """TrySoftRestoStep: equal primal and dual step lengths; accepted if the original criterion holds at alpha_test = 0
            or the averaged primal-dual error shrinks by soft_resto_pderror_reduction_factor."""
            al = min(a_max_of(stp), a_dual_of(stp))
            t = trial(al, stp)
            if not (math.isfinite(t["phi"]) and math.isfinite(t["theta"])):
                return None, al, False
            t.update(y=y + al * stp["y"], zl=zl + al * stp["zl"], zu=zu + al * stp["zu"])
            if acceptable(0.0, t["phi"], t["theta"]):
                return t, al, True
            t["grad"], t["J"] = derivs(t["x"])
            if pd_error(t, mu) <= (1.0 - 1e-4) * pd_error(P, mu):
                counters["soft_resto_steps"] += 1
                return t, al, False
            return None, al, False

        dvm = step["v"]
        vcur = np.concatenate([x, s])
        tiny = bool((np.abs(dvm) / (1.0 + np.abs(vcur)) <= 10.0 * EPS).all() and cur_theta <= 1e-4)
        if in_watchdog and tiny:       # StopWatchDog here would re-enter the line search at the stored point WITH second-order
            # corrections, i.e. with a fresh factorisation there: rare (never seen), not restated
            return dict(status="needs_globalisation", iters=it, why="tiny step inside the watchdog", counters=counters)
        if not in_watchdog and not tiny and not in_soft and wd_short >= 10:     # StartWatchDog (watchdog_shortened_iter_trigger)
            in_watchdog, wd_trial = True, 0
            wd = dict(P=P, step=step, dw=dw, ref=ref, alpha_test=a_max_of(step))
            counters["watchdog_starts"] += 1
            if events is not None:
                events.append((it, "watchdog_start"))
        got, a_pr, used, shortened, dual_done = None, 0.0, step, 0, False
        if tiny:
            counters["tiny_steps"] += 1
            a_pr = a_max_of(step)
            got = trial(a_pr, step); n_trials += 1
            if tiny_last:
                tiny_flag = True
            tiny_last = bool(np.abs(step["y"]).max() < 1e-2)                     # tiny_step_y_tol
        else:
            tiny_last = False
            if in_soft:
                soft_cnt += 1
                if soft_cnt <= 10:                                               # max_soft_resto_iters
                    got, a_pr, sat = soft_step(step)
                    dual_done = got is not None
                    if got is not None and sat:
                        in_soft, soft_cnt, dual_done, a_pr = False, 0, False, 0.0     # IPOPT then repeats the dual step with alpha_primal = 0
            else:
                skip_first = False
                while True:
                    got, a_pr, used, shortened = backtrack(step, P, skip_first)
                    if not in_watchdog:
                        break
                    if got is not None:
                        in_watchdog = False; break
                    wd_trial += 1
                    if wd_trial > 3:                                             # watchdog_trial_iter_max: back to the stored point
                        P, step, dw, in_watchdog, wd_short = wd["P"], wd["step"], wd["dw"], False, 0
                        x, s, y, zl, zu, f, g, grad, J = (P[k] for k in ("x", "s", "y", "zl", "zu", "f", "g", "grad", "J"))
                        cres = g - s
                        ref = wd["ref"]
                        skip_first = True
                        continue
                    got = trial(a_pr, step)                                     # accept the full step without test
                    break
        if got is None:
            if not in_soft:                                                      # the current direction as a soft restoration step
                augment_filter()
                got, a_pr, sat = soft_step(step)
                if got is not None:
                    dual_done = True
                    if sat:
                        dual_done, a_pr = False, 0.0
                    else:
                        in_soft = True
            if got is None:
                return dict(status="needs_restoration", iters=it, why="line search", counters=counters)
        f_before = P["f"]
        if not dual_done:
            a_du = a_dual_of(used)
            got.update(y=y + a_pr * used["y"], zl=zl + a_du * used["zl"], zu=zu + a_du * used["zu"])
            wd_short = 0 if shortened == 0 else wd_short + 1
        else:
            a_du = a_pr
        # ---- AcceptTrialPoint: kappa_sigma safeguard (the slack safeguard is not restated: stop if it would act)
        a, c = slacks(got["x"], got["s"])
        if (a[LO] < EPS * min(1.0, mu)).any() or (c[UP] < EPS * min(1.0, mu)).any():
            return dict(status="needs_slack_safeguard", iters=it, counters=counters)
        got["zl"] = np.where(LO, np.maximum(np.minimum(got["zl"], kappa_sigma * mu / a), mu / (kappa_sigma * a)), 0.0)
        got["zu"] = np.where(UP, np.maximum(np.minimum(got["zu"], kappa_sigma * mu / c), mu / (kappa_sigma * c)), 0.0)
        got["grad"], got["J"] = derivs(got["x"])
        P = {k: got[k] for k in ("x", "s", "y", "zl", "zu", "f", "g", "grad", "J")}
        if log is not None:
            log.append((mu, f_before / df, pr0, du0, dw, a_pr, a_du, n_trials))
        it += 1
    xo = np.minimum(np.maximum(P["x"], lbx), ubx)                # honor_original_bounds
    return dict(status=status, iters=it, x=xo, f=prob.fg(xo)[0], lam_x=(P["zu"][:n] - P["zl"][:n]) / df, lam_g=P["y"] * dc / df,
                counters=counters)
