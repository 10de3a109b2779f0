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
barrier update, fraction to the boundary, inertia correction, the filter line search with second-order corrections,
the kappa_sigma multiplier safeguard, the scaled-error convergence test.  NOT restated: watchdog, soft restoration,
restoration phase, tiny steps -- a solve that needs one of them stops with status "needs_globalisation", and the test
(tests/test_oracle_solve.py::test_fullspace_ipm_reproduces_oracle_iterates) uses instances on which the C++ oracle's
own counters say that none of them ran.  Only tests/ may import this module.
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


def solve(prob: Problem, x0, lbx, ubx, lbg, ubg, *, max_iter=100, tol=1e-8, log=None):
    """One cold-started solve.  Returns dict(status, iters, x, f, lam_x, lam_g); log (a list) receives one tuple
    (mu, f, inf_pr, inf_du, dw, alpha_pr, alpha_du, trial points) per iteration, as IPOPT's iteration output would."""
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

    filt: list[tuple[float, float]] = []
    theta_max = theta_min = None
    dw_last = 0.0
    tau = max(tau_min, 1.0 - mu)
    it = 0
    status = None
    while status is None:
        if converged():
            status = "Solve_Succeeded"; break
        if it >= max_iter:
            status = "Maximum_Iterations_Exceeded"; break
        # ---- monotone barrier update (MonotoneMuUpdate, mu_allow_fast_monotone_decrease)
        mu_floor = min(tol, df * compl_inf_tol) / (kappa_eps + 1.0)
        while errors(x, s, y, zl, zu, g, grad, J, mu)[3] <= kappa_eps * mu:
            new = max(mu_floor, min(kappa_mu * mu, mu ** theta_mu))
            if new == mu:
                break
            mu = new; tau = max(tau_min, 1.0 - mu); filt = []
        du0, pr0, _, _ = errors(x, s, y, zl, zu, g, grad, J, 0.0)
        # ---- search direction from the full-space system, inertia from the spectrum
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
                return dict(status="needs_globalisation", iters=it, why="inertia")
        if dw > 0:
            dw_last = dw

        def direction(cres):
            sol = np.linalg.solve(K, -np.concatenate([rhs1, cres]))
            dv, dy_ = sol[:n + m], sol[n + m:]
            dzl = np.where(LO, (mu - zl * dv) / a - zl, 0.0)
            dzu = np.where(UP, (mu + zu * dv) / c - zu, 0.0)
            return dv, dy_, dzl, dzu

        def ftb(vals, steps, mask):
            neg = mask & (steps < 0)
            return min(1.0, (-tau * vals[neg] / steps[neg]).min(initial=1.0))

        cres = g - s
        dv, dy, dzl, dzu = direction(cres)
        a_max = min(ftb(a, dv, LO), ftb(c, -dv, UP))
        # ---- filter line search (BacktrackingLineSearch + FilterLSAcceptor)
        theta = np.abs(cres).sum()
        phi = barrier(x, s, f, mu)
        gbd = float(gphi @ dv)
        if theta_max is None:
            theta_max = 1e4 * max(1.0, theta); theta_min = 1e-4 * max(1.0, theta)

        def ftype(al):
            if theta == 0.0 and 0.0 < gbd < 100.0 * EPS:
                return True
            return gbd < 0 and al * (-gbd) ** s_phi > delta * theta ** s_theta

        def armijo(al, phi_t):
            return _cmp_le(phi_t - phi, eta_phi * al * gbd, phi)

        def acceptable(al, phi_t, th_t):
            if not (math.isfinite(phi_t) and math.isfinite(th_t)) or th_t > theta_max:
                return False
            if ftype(al) and theta <= theta_min:
                ok = armijo(al, phi_t)
            else:
                if phi_t > phi:
                    bas = math.log10(abs(phi)) if abs(phi) > 10.0 else 1.0
                    if math.log10(phi_t - phi) > obj_max_inc + bas:
                        return False
                ok = _cmp_le(th_t, (1 - g_theta) * theta, theta) or _cmp_le(phi_t - phi, -g_phi * theta, phi)
            return ok and all(_cmp_le(phi_t, fb, fb) or _cmp_le(th_t, ft, ft) for fb, ft in filt)

        a_min = g_theta
        if gbd < 0:
            a_min = min(g_theta, g_phi * theta / (-gbd))
            if theta <= theta_min:
                a_min = min(a_min, delta * theta ** s_theta / (-gbd) ** s_phi)
        a_min *= a_min_frac

        def trial(al, step):
            xt = x + al * step[:n]; st_ = s + al * step[n:]
            ft, gt = fg(xt)
            return xt, st_, ft, gt, np.abs(gt - st_).sum(), barrier(xt, st_, ft, mu)

        alpha, n_trials, shortened, accepted = a_max, 0, 0, None
        step = (dv, dy, dzl, dzu)
        while alpha > a_min or shortened == 0:
            xt, st_, ft, gt, th_t, phi_t = trial(alpha, step[0]); n_trials += 1
            if acceptable(alpha, phi_t, th_t):
                accepted = (alpha, xt, st_, ft, gt, phi_t); break
            if alpha == a_max and theta <= th_t and math.isfinite(phi_t):
                # second-order correction: same matrix, constraint residual alpha * c + c(trial), accumulated
                csoc, a_soc, th_old, cnt = cres.copy(), alpha, 0.0, 0
                th_soc, g_soc, s_soc = th_t, gt, st_
                while cnt < max_soc and accepted is None and (cnt == 0 or th_soc <= kappa_soc * th_old):
                    th_old = th_soc
                    csoc = a_soc * csoc + (g_soc - s_soc)
                    stp = direction(csoc)
                    a_soc = min(ftb(a, stp[0], LO), ftb(c, -stp[0], UP))
                    xt, st_, ft, gt, th_t2, phi_t2 = trial(a_soc, stp[0]); n_trials += 1
                    if not math.isfinite(phi_t2):
                        break
                    if acceptable(alpha, phi_t2, th_t2):
                        accepted = (a_soc, xt, st_, ft, gt, phi_t2); step = stp
                    else:
                        cnt += 1; th_soc, g_soc, s_soc = th_t2, gt, st_
                if accepted is not None:
                    break
            alpha *= 0.5; shortened += 1
        if accepted is None:
            return dict(status="needs_globalisation", iters=it, why="line search")
        a_pr, xt, st_, ft, gt, phi_t = accepted
        f_before = f
        if not (ftype(alpha) and armijo(alpha, phi_t)):       # not an Armijo-accepted f-type step: augment the filter
            ne = (phi - g_phi * theta, (1 - g_theta) * theta)
            filt = [e for e in filt if not (e[0] >= ne[0] and e[1] >= ne[1])] + [ne]
        dv, dy, dzl, dzu = step
        a_du = min(ftb(zl, dzl, LO), ftb(zu, dzu, UP))
        y = y + a_pr * dy
        zl = zl + a_du * dzl; zu = zu + a_du * dzu
        x, s, f, g = xt, st_, ft, gt
        # kappa_sigma safeguard (IpoptAlgorithm::correct_bound_multiplier)
        a, c = slacks(x, s)
        zl = np.where(LO, np.maximum(np.minimum(zl, kappa_sigma * mu / a), mu / (kappa_sigma * a)), 0.0)
        zu = np.where(UP, np.maximum(np.minimum(zu, kappa_sigma * mu / c), mu / (kappa_sigma * c)), 0.0)
        grad, J = derivs(x)
        if log is not None:
            log.append((mu, f_before / df, pr0, du0, dw, a_pr, a_du, n_trials))
        it += 1
    xo = np.minimum(np.maximum(x, lbx), ubx)                     # honor_original_bounds
    return dict(status=status, iters=it, x=xo, f=prob.fg(xo)[0], lam_x=(zu[:n] - zl[:n]) / df, lam_g=y * dc / df)
