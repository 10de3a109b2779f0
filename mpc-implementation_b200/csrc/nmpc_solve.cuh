// nmpc_solve.cuh -- one-warp-per-instance primal-dual interior-point solve (the hot path).
//
// Replaces  sol = solver(x0,lbx,ubx,lbg,ubg,p)  (Python/NMPC_TT.py:358-365; CasADi -> IPOPT -> MUMPS) for the
// reference's UAV target-tracking NLP.  Algorithm = IPOPT's (Waechter & Biegler 2006) with the scripts'
// options (NMPC_TT.py:257-265): slack form g(w) - s = 0, relaxed bounds, gradient-based scaling,
// monotone barrier, fraction-to-boundary, inertia correction, filter line search + second-order
// correction, scaled optimality-error termination.  The whole iteration loop runs inside the kernel.
#pragma once
#include "nmpc_device.cuh"

namespace nmpc {

struct Opt {
  int max_iter; int scaling; double tol;
  double dual_inf_tol = 1.0, constr_viol_tol = 1e-4, compl_inf_tol = 1e-4;
  double bound_relax = 1e-8, bound_push = 1e-2, bound_frac = 1e-2;
  double mu_init = 0.1, kappa_mu = 0.2, theta_mu = 1.5, kappa_eps = 10.0, tau_min = 0.99;
  double kappa_d = 1e-4, kappa_sigma = 1e10, s_max = 100.0;
  double max_grad = 100.0, scal_min = 1e-8;
  double constr_mult_init_max = 1e3;
  double dw_init = 1e-4, dw_min = 1e-20, dw_max = 1e20, dw_inc_first = 100.0, dw_inc = 8.0, dw_dec = 1.0 / 3.0;
  double gamma_theta = 1e-5, gamma_phi = 1e-8, eta_phi = 1e-8, s_theta = 1.1, s_phi = 2.3, delta = 1.0;
  double alpha_min_frac = 0.05, alpha_red = 0.5; int max_soc = 4; double kappa_soc = 0.99;
  double theta_max_fact = 1e4, theta_min_fact = 1e-4;
  double tiny_step_tol = 10.0 * 2.220446049250313e-16;
};

struct SolveArgs {
  Prob pr; Opt o;
  int B;
  const double *p, *x0, *lbx, *ubx, *lbg, *ubg, *obs;
  int obs_per_instance;
  double *x, *f, *g, *lam_x, *lam_g;
  int32_t *status, *iters;
  int* counter;                      // work queue
  unsigned long long* stats;         // [3]: factorizations, ls trials, soc accepted
  int ws_doubles;                    // per-warp workspace size
  double* dbg; int dbg_rows;         // optional per-iteration log [B][dbg_rows][8] (tests only)
};

__host__ __device__ inline int ws_size(int S, int R, int n_obs) {
  int n = A_NROW * R * S + LQ_N * S + RIC_N * S + 3 * n_obs + 2 * FILT_CAP;
  return (n + 1) & ~1;
}

constexpr double EPSM = 2.220446049250313e-16;
__device__ __forceinline__ bool cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * EPSM * fabs(bas); }
__device__ __forceinline__ bool is_lo(double b) { return b > -1e300; }
__device__ __forceinline__ bool is_hi(double b) { return b < 1e300; }

__device__ __forceinline__ double push_in(double v, double lo, double hi, double k1, double k2) {
  const bool hl = is_lo(lo), hu = is_hi(hi);
  if (hl && hu) {
    const double pl = fmin(k1 * fmax(1.0, fabs(lo)), k2 * (hi - lo));
    const double pu = fmin(k1 * fmax(1.0, fabs(hi)), k2 * (hi - lo));
    v = fmax(v, lo + pl); v = fmin(v, hi - pu);
  } else if (hl) v = fmax(v, lo + k1 * fmax(1.0, fabs(lo)));
  else if (hu) v = fmin(v, hi - k1 * fmax(1.0, fabs(hi)));
  return v;
}

// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void solve_instance(const SolveArgs& A, double* __restrict__ ws, int b, int lane) {
  const Prob& pr = A.pr; const Opt& o = A.o;
  const int N = pr.N, S = pr.S, R = pr.R, n_obs = pr.n_obs;
  const bool act = lane <= N, hasu = lane < N;
  const double T = pr.T;
  double* rows = ws;
  double* lq = rows + A_NROW * R * S;
  double* ric = lq + LQ_N * S;
  double* obs = ric + RIC_N * S;
  double* filt = obs + 3 * n_obs;
#define RW(arr, r) rows[((arr) * R + (r)) * S + lane]
#define LQ(e) lq[(e) * S + lane]
  const double NINF = -CUDART_INF, PINF = CUDART_INF;

  // ---------------- load instance ------------------------------------------------------------
  double X0[8], xt, yt;
  {
    const double* pp = A.p + (size_t)b * NPAR;
#pragma unroll
    for (int i = 0; i < 8; ++i) X0[i] = pp[i];
    xt = pp[8]; yt = pp[9];
  }
  double u[6], xL[6], xU[6], zL[6], zU[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) { u[i] = 0.0; xL[i] = NINF; xU[i] = PINF; zL[i] = 0.0; zU[i] = 0.0; }
  if (hasu) {
    const double* xx = A.x0 + (size_t)b * (NU * N) + NU * lane;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      u[i] = xx[i];
      const double lo = A.lbx[NU * lane + i], hi = A.ubx[NU * lane + i];
      xL[i] = lo > -1e19 ? lo - o.bound_relax * fmax(1.0, fabs(lo)) : NINF;
      xU[i] = hi < 1e19 ? hi + o.bound_relax * fmax(1.0, fabs(hi)) : PINF;
    }
  }
  {
    const double* ob = A.obs + (A.obs_per_instance ? (size_t)b * 3 * n_obs : 0);
    for (int i = lane; i < 3 * n_obs; i += 32) obs[i] = ob[i];
  }
  __syncwarp();

  Stage st;
  double df = 1.0;
  if (act) for (int r = 0; r < R; ++r) RW(A_DC, r) = 1.0;

  // stage-local helpers ------------------------------------------------------------------------
  // T-scaled dynamics Jacobian entries of this lane's stage
  double e03, e13, e23, e04, e14;
  auto dyn_entries = [&](const Stage& s_, const double* u_) {
    const double tv = hasu ? T * u_[0] : 0.0;
    e03 = -tv * s_.cps * s_.sth; e13 = -tv * s_.sps * s_.sth; e23 = tv * s_.cth;
    e04 = -tv * s_.sps * s_.cth; e14 = tv * s_.cps * s_.cth;
  };
  // adjoint recursion lam_k = a_k + A_k^T lam_{k+1} by suffix scans
  auto adjoint = [&](const double* a, double* lam, double* lamn) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i == 3 || i == 4) continue;
      lam[i] = rscan_incl(act ? a[i] : 0.0, lane); lamn[i] = shfl_next(lam[i], lane);
    }
    const double b3 = (act ? a[3] : 0.0) + e03 * lamn[0] + e13 * lamn[1] + e23 * lamn[2];
    const double b4 = (act ? a[4] : 0.0) + e04 * lamn[0] + e14 * lamn[1];
    lam[3] = rscan_incl(b3, lane); lamn[3] = shfl_next(lam[3], lane);
    lam[4] = rscan_incl(b4, lane); lamn[4] = shfl_next(lam[4], lane);
  };
  // scaled constraint values of stage `lane` into row array `arr`
  auto eval_g = [&](const Stage& s_, int arr) {
    if (act) {
#pragma unroll
      for (int r = 0; r < 5; ++r) RW(arr, r) = RW(A_DC, r) * s_.X[box_state(r)];
      for (int jn = 0; jn < n_obs; ++jn) {
        const double dx_ = s_.X[0] - obs[3 * jn], dy_ = s_.X[1] - obs[3 * jn + 1];
        RW(arr, 5 + jn) = RW(A_DC, 5 + jn) * (obs[3 * jn + 2] - sqrt(dx_ * dx_ + dy_ * dy_));
      }
    }
  };
  auto cost_sum = [&](const Stage& s_) -> double {
    const double l = hasu ? stage_cost(pr, s_.X, xt, yt) : 0.0;
    return df * warp_sum(l);
  };

  // ---------------- gradient-based scaling at the user's starting point ----------------------
  if (o.scaling) {
    rollout(pr, X0, u, lane, st);
    dyn_entries(st, u);
    double gl[6], Hl[21], a[8], lam[8], lamn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.0;
    if (hasu && lane >= 1) {
      stage_cost_d2(pr, st.X, xt, yt, gl, Hl);
#pragma unroll
      for (int v = 0; v < 6; ++v) a[cost_state(v)] = gl[v];
    }
    adjoint(a, lam, lamn);
    double gmax = 0.0;
    if (hasu) {
      gmax = fabs(T * (st.cps * st.cth * lamn[0] + st.sps * st.cth * lamn[1] + st.sth * lamn[2]));
#pragma unroll
      for (int r = 1; r < 6; ++r) gmax = fmax(gmax, fabs(T * lamn[r + 2]));
    }
    gmax = warp_max(gmax);
    // row maxima of the Jacobian: |d g_{k,i} / d u_j| for j < k
    const double d0 = st.cps * st.cth, d1 = st.sps * st.cth, d2 = st.sth;
    const double c30 = scan_excl(e03, lane), c31 = scan_excl(e13, lane), c32 = scan_excl(e23, lane);
    const double c40 = scan_excl(e04, lane), c41 = scan_excl(e14, lane);
    double zmax = 0.0;
    // scratch in row arrays that are not live yet: A_GT = row max, A_DS / A_DS2 = obstacle normal
    if (act) for (int jn = 0; jn < n_obs; ++jn) {
      const double dx_ = st.X[0] - obs[3 * jn], dy_ = st.X[1] - obs[3 * jn + 1];
      const double iD = 1.0 / sqrt(dx_ * dx_ + dy_ * dy_);
      RW(A_GT, 5 + jn) = 0.0; RW(A_DS, 5 + jn) = dx_ * iD; RW(A_DS2, 5 + jn) = dy_ * iD;
    }
    for (int jj = 0; jj < N; ++jj) {
      const double dj0 = __shfl_sync(FULL, d0, jj), dj1 = __shfl_sync(FULL, d1, jj), dj2 = __shfl_sync(FULL, d2, jj);
      const double a30 = __shfl_sync(FULL, c30, jj + 1), a31 = __shfl_sync(FULL, c31, jj + 1), a32 = __shfl_sync(FULL, c32, jj + 1);
      const double a40 = __shfl_sync(FULL, c40, jj + 1), a41 = __shfl_sync(FULL, c41, jj + 1);
      if (act && jj < lane) {
        const double pv0 = T * dj0, pv1 = T * dj1, pv2 = T * dj2;
        const double pt0 = T * (c30 - a30), pt1 = T * (c31 - a31), pt2 = T * (c32 - a32);
        const double pp0 = T * (c40 - a40), pp1 = T * (c41 - a41);
        zmax = fmax(zmax, fmax(fabs(pv2), fabs(pt2)));
        for (int jn = 0; jn < n_obs; ++jn) {
          const double nx = RW(A_DS, 5 + jn), ny = RW(A_DS2, 5 + jn);
          const double m1 = fabs(nx * pv0 + ny * pv1), m2 = fabs(nx * pt0 + ny * pt1), m3 = fabs(nx * pp0 + ny * pp1);
          RW(A_GT, 5 + jn) = fmax(RW(A_GT, 5 + jn), fmax(m1, fmax(m2, m3)));
        }
      }
    }
    if (act) {
      auto sc = [&](double m) { return m > o.max_grad ? fmax(o.scal_min, o.max_grad / m) : 1.0; };
      RW(A_DC, 0) = sc(zmax);
      const double lin = lane >= 1 ? T : 0.0;
#pragma unroll
      for (int r = 1; r < 5; ++r) RW(A_DC, r) = sc(lin);
      for (int jn = 0; jn < n_obs; ++jn) RW(A_DC, 5 + jn) = sc(RW(A_GT, 5 + jn));
    }
    df = gmax > o.max_grad ? fmax(o.scal_min, o.max_grad / gmax) : 1.0;
  }
  // ---------------- scaled + relaxed constraint bounds ---------------------------------------
  if (act) {
    for (int r = 0; r < R; ++r) {
      const double lo = A.lbg[lane * R + r], hi = A.ubg[lane * R + r], dc = RW(A_DC, r);
      double l2 = NINF, h2 = PINF;
      if (lo > -1e19) { l2 = dc * lo; l2 -= o.bound_relax * fmax(1.0, fabs(l2)); }
      if (hi < 1e19) { h2 = dc * hi; h2 += o.bound_relax * fmax(1.0, fabs(h2)); }
      RW(A_DL, r) = l2; RW(A_DU, r) = h2;
    }
  }
  // ---------------- starting point -------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    if (hasu) { u[i] = push_in(u[i], xL[i], xU[i], o.bound_push, o.bound_frac); zL[i] = is_lo(xL[i]) ? 1.0 : 0.0; zU[i] = is_hi(xU[i]) ? 1.0 : 0.0; }
  }
  rollout(pr, X0, u, lane, st);
  double f = cost_sum(st);
  eval_g(st, A_G);
  if (act) {
    for (int r = 0; r < R; ++r) {
      const double lo = RW(A_DL, r), hi = RW(A_DU, r);
      RW(A_S, r) = push_in(RW(A_G, r), lo, hi, o.bound_push, o.bound_frac);
      RW(A_VL, r) = is_lo(lo) ? 1.0 : 0.0; RW(A_VU, r) = is_hi(hi) ? 1.0 : 0.0;
      RW(A_Y, r) = 0.0;
    }
  }
  double mu = o.mu_init, tau = fmax(o.tau_min, 1.0 - mu);
  const double mu_floor = fmin(o.tol, o.compl_inf_tol) / (o.kappa_eps + 1.0);
  double theta_max, theta_min;
  {
    double th = 0.0;
    if (act) for (int r = 0; r < R; ++r) th += fabs(RW(A_G, r) - RW(A_S, r));
    th = warp_sum(th);
    theta_max = o.theta_max_fact * fmax(1.0, th); theta_min = o.theta_min_fact * fmax(1.0, th);
  }
  int nfilt = 0;
  double dw_last = 0.0;
  unsigned long long n_fact = 0, n_ls = 0, n_soc = 0;

  // barrier pieces of the current point: LB = sum log(slack), DT = sum of one-sided slacks
  auto barrier_parts = [&](const double* u_, int sarr_is_trial, double alpha, int dsarr, double& LB, double& DT) {
    // slack values: s (current) or s + alpha*ds (trial)
    double lb = 0.0, dt = 0.0;
    if (hasu) {
      double prod = 1.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const bool hl = is_lo(xL[i]), hu = is_hi(xU[i]);
        if (hl) prod *= (u_[i] - xL[i]);
        if (hu) prod *= (xU[i] - u_[i]);
        if (hl && !hu) dt += u_[i] - xL[i];
        if (hu && !hl) dt += xU[i] - u_[i];
        if (i == 2 || i == 5) { lb += log(prod); prod = 1.0; }
      }
    }
    if (act) {
      double prod = 1.0; int cnt = 0;
      for (int r = 0; r < R; ++r) {
        const double sv = sarr_is_trial ? RW(A_S, r) + alpha * RW(dsarr, r) : RW(A_S, r);
        const double lo = RW(A_DL, r), hi = RW(A_DU, r);
        const bool hl = is_lo(lo), hu = is_hi(hi);
        if (hl) { prod *= (sv - lo); ++cnt; }
        if (hu) { prod *= (hi - sv); ++cnt; }
        if (hl && !hu) dt += sv - lo;
        if (hu && !hl) dt += hi - sv;
        if (cnt >= 4) { lb += log(prod); prod = 1.0; cnt = 0; }
      }
      if (cnt) lb += log(prod);
    }
    LB = warp_sum(lb); DT = warp_sum(dt);
  };
  double LB, DT;
  barrier_parts(u, 0, 0.0, A_DS, LB, DT);

  // ---- derivative evaluation at the current point: fills Q (Hessian of the Lagrangian in stage form),
  //      dynamics data, and returns the cost gradient gl (scaled by df) and the adjoint of the Lagrangian.
  double gl[6], lam[8], lamn[8];
  auto eval_derivs = [&](bool ls_mode) {
    // ls_mode: least-squares multiplier system (W = 0, Sigma = I)
    dyn_entries(st, u);
    double Hl[21];
#pragma unroll
    for (int v = 0; v < 6; ++v) gl[v] = 0.0;
#pragma unroll
    for (int e = 0; e < 21; ++e) Hl[e] = 0.0;
    if (hasu && lane >= 1) {
      stage_cost_d2(pr, st.X, xt, yt, gl, Hl);
#pragma unroll
      for (int v = 0; v < 6; ++v) gl[v] *= df;
    }
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.0;
    if (act) {
#pragma unroll
      for (int e = 0; e < 36; ++e) LQ(LQ_Q + e) = 0.0;
      if (!ls_mode) {
#pragma unroll
        for (int uu = 0; uu < 6; ++uu)
#pragma unroll
          for (int v = 0; v <= uu; ++v) LQ(LQ_Q + tri(cost_state(uu), cost_state(v))) = df * Hl[tri(uu, v)];
      }
#pragma unroll
      for (int v = 0; v < 6; ++v) a[cost_state(v)] = gl[v];
      // box rows
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        const int i = box_state(r);
        const double dc = RW(A_DC, r), sv = RW(A_S, r), lo = RW(A_DL, r), hi = RW(A_DU, r);
        double sig = 1.0;
        if (!ls_mode) sig = (is_lo(lo) ? RW(A_VL, r) / (sv - lo) : 0.0) + (is_hi(hi) ? RW(A_VU, r) / (hi - sv) : 0.0);
        LQ(LQ_Q + tri(i, i)) += dc * dc * sig;
        LQ(LQ_DG + r) = dc * dc;
        a[i] += dc * RW(A_Y, r);
      }
      // obstacle rows
      double nn0 = 0.0, nn1 = 0.0, nn2 = 0.0, q00 = 0.0, q01 = 0.0, q11 = 0.0;
      for (int jn = 0; jn < n_obs; ++jn) {
        const int r = 5 + jn;
        const double dx_ = st.X[0] - obs[3 * jn], dy_ = st.X[1] - obs[3 * jn + 1];
        const double D = sqrt(dx_ * dx_ + dy_ * dy_), iD = 1.0 / D, nx = dx_ * iD, ny = dy_ * iD;
        const double dc = RW(A_DC, r), sv = RW(A_S, r), lo = RW(A_DL, r), hi = RW(A_DU, r), yv = RW(A_Y, r);
        double sig = 1.0;
        if (!ls_mode) sig = (is_lo(lo) ? RW(A_VL, r) / (sv - lo) : 0.0) + (is_hi(hi) ? RW(A_VU, r) / (hi - sv) : 0.0);
        const double w = dc * dc;
        nn0 += w * nx * nx; nn1 += w * nx * ny; nn2 += w * ny * ny;
        const double cur = ls_mode ? 0.0 : -yv * dc * iD;   // y * d2h,  d2h = -(I - n n^T)/D
        q00 += w * sig * nx * nx + cur * (1.0 - nx * nx);
        q01 += w * sig * nx * ny + cur * (-nx * ny);
        q11 += w * sig * ny * ny + cur * (1.0 - ny * ny);
        a[0] += -dc * yv * nx; a[1] += -dc * yv * ny;
      }
      LQ(LQ_Q + tri(0, 0)) += q00; LQ(LQ_Q + tri(1, 0)) += q01; LQ(LQ_Q + tri(1, 1)) += q11;
      LQ(LQ_NN + 0) = nn0; LQ(LQ_NN + 1) = nn1; LQ(LQ_NN + 2) = nn2;
      LQ(LQ_DD + 0) = st.cps * st.cth; LQ(LQ_DD + 1) = st.sps * st.cth; LQ(LQ_DD + 2) = st.sth;
      LQ(LQ_EE + 0) = e03; LQ(LQ_EE + 1) = e13; LQ(LQ_EE + 2) = e23; LQ(LQ_EE + 3) = e04; LQ(LQ_EE + 4) = e14;
    }
    adjoint(a, lam, lamn);
    if (act) {
      double svt = 0.0, svp = 0.0;
      if (hasu && !ls_mode) {
        // curvature of T*v*d(theta,psi) weighted by the next-stage adjoint
        const double L0 = T * lamn[0], L1 = T * lamn[1], L2 = T * lamn[2], v = u[0];
        const double cc = st.cps * st.cth, sc = st.sps * st.cth, cs = st.cps * st.sth, ss = st.sps * st.sth;
        LQ(LQ_Q + tri(3, 3)) += -v * (L0 * cc + L1 * sc + L2 * st.sth);
        LQ(LQ_Q + tri(4, 4)) += -v * (L0 * cc + L1 * sc);
        LQ(LQ_Q + tri(4, 3)) += v * (L0 * ss - L1 * cs);
        svt = -L0 * cs - L1 * ss + L2 * st.cth;
        svp = -L0 * sc + L1 * cc;
      }
      LQ(LQ_SV + 0) = svt; LQ(LQ_SV + 1) = svp;
    }
  };

  // Newton right-hand side (q, r, Sigma_x) for barrier parameter mu into the LQ arrays.
  auto build_rhs = [&](bool ls_mode) {
    if (act) {
      double q[8], qd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { q[i] = 0.0; qd[i] = 0.0; }
#pragma unroll
      for (int v = 0; v < 6; ++v) q[cost_state(v)] = gl[v];
      for (int r = 0; r < R; ++r) {
        const double dc = RW(A_DC, r), sv = RW(A_S, r), lo = RW(A_DL, r), hi = RW(A_DU, r);
        const bool hl = is_lo(lo), hu = is_hi(hi);
        double yh, cd;   // yh = y + Sigma_s c + rs (delta_w-free part), cd = c
        if (ls_mode) { yh = -RW(A_VL, r) + RW(A_VU, r); cd = 0.0; }
        else {
          const double sig = (hl ? RW(A_VL, r) / (sv - lo) : 0.0) + (hu ? RW(A_VU, r) / (hi - sv) : 0.0);
          const double c = RW(A_G, r) - sv;
          double bg = (hl ? -mu / (sv - lo) : 0.0) + (hu ? mu / (hi - sv) : 0.0);
          if (hl && !hu) bg += o.kappa_d * mu; if (hu && !hl) bg -= o.kappa_d * mu;
          yh = sig * c + bg; cd = c;
        }
        if (r < 5) { const int i = box_state(r); q[i] += dc * yh; qd[i] += dc * cd; }
        else {
          const int jn = r - 5;
          const double dx_ = st.X[0] - obs[3 * jn], dy_ = st.X[1] - obs[3 * jn + 1];
          const double iD = 1.0 / sqrt(dx_ * dx_ + dy_ * dy_), nx = dx_ * iD, ny = dy_ * iD;
          q[0] += -dc * yh * nx; q[1] += -dc * yh * ny; qd[0] += -dc * cd * nx; qd[1] += -dc * cd * ny;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { LQ(LQ_QV + i) = q[i]; LQ(LQ_QD + i) = qd[i]; }
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double sig = 1.0, rr = 0.0;
        if (hasu) {
          const bool hl = is_lo(xL[i]), hu = is_hi(xU[i]);
          if (ls_mode) rr = -zL[i] + zU[i];
          else {
            sig = (hl ? zL[i] / (u[i] - xL[i]) : 0.0) + (hu ? zU[i] / (xU[i] - u[i]) : 0.0);
            rr = (hl ? -mu / (u[i] - xL[i]) : 0.0) + (hu ? mu / (xU[i] - u[i]) : 0.0);
            if (hl && !hu) rr += o.kappa_d * mu; if (hu && !hl) rr -= o.kappa_d * mu;
          }
        }
        LQ(LQ_RD + i) = sig; LQ(LQ_RV + i) = rr;
      }
    }
    __syncwarp();
  };

  // ---------------- least-squares multiplier start -------------------------------------------
  double mydx[8], mydu[6];
  {
    eval_derivs(true);
    build_rhs(true);
    ++n_fact;
    const bool ok = riccati_factor(pr, lq, ric, 0.0, lane);
    double ymax = 0.0;
    if (ok) {
      riccati_forward(pr, lq, ric, false, lane, mydx, mydu);
      if (act) {
        for (int r = 0; r < R; ++r) {
          const double dc = RW(A_DC, r);
          double gd;
          if (r < 5) gd = dc * mydx[box_state(r)];
          else {
            const int jn = r - 5;
            const double dx_ = st.X[0] - obs[3 * jn], dy_ = st.X[1] - obs[3 * jn + 1];
            const double iD = 1.0 / sqrt(dx_ * dx_ + dy_ * dy_);
            gd = -dc * (dx_ * iD * mydx[0] + dy_ * iD * mydx[1]);
          }
          const double yv = gd + (-RW(A_VL, r) + RW(A_VU, r));
          RW(A_Y, r) = yv; ymax = fmax(ymax, fabs(yv));
        }
      }
      ymax = warp_max(ymax);
    }
    if (!ok || !(ymax <= o.constr_mult_init_max)) { if (act) for (int r = 0; r < R; ++r) RW(A_Y, r) = 0.0; }
    __syncwarp();
  }

  // ---------------- main loop ------------------------------------------------------------------
  int iter = 0, status = NMPC_MAXITER_EXCEEDED, tiny_count = 0; bool tiny_flag = false;
  for (;;) {
    eval_derivs(false);
    // ---- optimality error
    double du_l = 0.0, pr_l = 0.0, sumy = 0.0, sumz = 0.0, viol = 0.0; int nz = 0;
    double glx[6];
    if (hasu) {
      glx[0] = T * (st.cps * st.cth * lamn[0] + st.sps * st.cth * lamn[1] + st.sth * lamn[2]);
#pragma unroll
      for (int r = 1; r < 6; ++r) glx[r] = T * lamn[r + 2];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        du_l = fmax(du_l, fabs(glx[i] - zL[i] + zU[i]));
        if (is_lo(xL[i])) { sumz += fabs(zL[i]); ++nz; }
        if (is_hi(xU[i])) { sumz += fabs(zU[i]); ++nz; }
      }
    }
    if (act) {
      for (int r = 0; r < R; ++r) {
        const double yv = RW(A_Y, r), gv = RW(A_G, r), sv = RW(A_S, r), lo = RW(A_DL, r), hi = RW(A_DU, r), dc = RW(A_DC, r);
        du_l = fmax(du_l, fabs(-yv - RW(A_VL, r) + RW(A_VU, r)));
        pr_l = fmax(pr_l, fabs(gv - sv));
        sumy += fabs(yv);
        if (is_lo(lo)) { sumz += fabs(RW(A_VL, r)); ++nz; }
        if (is_hi(hi)) { sumz += fabs(RW(A_VU, r)); ++nz; }
        const double gu = gv / dc, lo_o = A.lbg[lane * R + r], hi_o = A.ubg[lane * R + r];
        if (lo_o > -1e19) viol = fmax(viol, lo_o - gu);
        if (hi_o < 1e19) viol = fmax(viol, gu - hi_o);
      }
    }
    const double du_inf = warp_max(du_l), pr_inf = warp_max(pr_l);
    sumy = warp_sum(sumy); sumz = warp_sum(sumz); viol = warp_max(viol);
    const int nzt = __reduce_add_sync(FULL, nz);
    const int mtot = R * S;
    const double sd = fmax(o.s_max, (sumy + sumz) / (double)max(1, mtot + nzt)) / o.s_max;
    const double sc = fmax(o.s_max, sumz / (double)max(1, nzt)) / o.s_max;
    auto compl_err = [&](double mu_) {
      double co = 0.0;
      if (hasu) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (is_lo(xL[i])) co = fmax(co, fabs((u[i] - xL[i]) * zL[i] - mu_));
          if (is_hi(xU[i])) co = fmax(co, fabs((xU[i] - u[i]) * zU[i] - mu_));
        }
      }
      if (act) for (int r = 0; r < R; ++r) {
        const double sv = RW(A_S, r), lo = RW(A_DL, r), hi = RW(A_DU, r);
        if (is_lo(lo)) co = fmax(co, fabs((sv - lo) * RW(A_VL, r) - mu_));
        if (is_hi(hi)) co = fmax(co, fabs((hi - sv) * RW(A_VU, r) - mu_));
      }
      return warp_max(co);
    };
    const double co0 = compl_err(0.0);
    const double E0 = fmax(du_inf / sd, fmax(pr_inf, co0 / sc));
    if (!isfinite(E0) || !isfinite(f)) { status = NMPC_INVALID_NUMBER; break; }
    if (E0 <= o.tol && du_inf / df <= o.dual_inf_tol && viol <= o.constr_viol_tol && co0 / df <= o.compl_inf_tol) {
      status = NMPC_SOLVE_SUCCEEDED; break;
    }
    if (iter >= o.max_iter) { status = NMPC_MAXITER_EXCEEDED; break; }
    // ---- barrier parameter
    {
      double Emu = fmax(du_inf / sd, fmax(pr_inf, compl_err(mu) / sc));
      while ((Emu <= o.kappa_eps * mu || tiny_flag) && mu > mu_floor) {
        mu = fmax(mu_floor, fmin(o.kappa_mu * mu, pow(mu, o.theta_mu)));
        tau = fmax(o.tau_min, 1.0 - mu); nfilt = 0; tiny_flag = false;
        Emu = fmax(du_inf / sd, fmax(pr_inf, compl_err(mu) / sc));
      }
      if (tiny_flag && mu <= mu_floor) { status = NMPC_STEP_TOO_SMALL; break; }
    }
    // ---- search direction with inertia correction
    const unsigned long long ls_before = n_ls;
    build_rhs(false);
    double dw = 0.0; bool ok = false;
    for (;;) {
      ++n_fact;
      ok = riccati_factor(pr, lq, ric, dw, lane);
      if (ok) break;
      if (dw == 0.0) dw = (dw_last == 0.0) ? o.dw_init : fmax(o.dw_min, dw_last * o.dw_dec);
      else dw = (dw_last == 0.0 || 1e5 * dw_last < dw) ? o.dw_inc_first * dw : o.dw_inc * dw;
      if (dw > o.dw_max) break;
    }
    if (!ok) { status = NMPC_PERTURBATION_FAILED; break; }
    if (dw > 0.0) dw_last = dw;
    riccati_forward(pr, lq, ric, false, lane, mydx, mydu);

    // per-row step ds = G dx + c into array `dsarr` for residual array c = (carr ? csoc : g - s)
    auto row_steps = [&](const double* dx_, int dsarr, bool soc) {
      if (act) for (int r = 0; r < R; ++r) {
        const double dc = RW(A_DC, r);
        double gd;
        if (r < 5) gd = dc * dx_[box_state(r)];
        else {
          const int jn = r - 5;
          const double ddx = st.X[0] - obs[3 * jn], ddy = st.X[1] - obs[3 * jn + 1];
          const double iD = 1.0 / sqrt(ddx * ddx + ddy * ddy);
          gd = -dc * (ddx * iD * dx_[0] + ddy * iD * dx_[1]);
        }
        const double c = soc ? RW(A_CSOC, r) : RW(A_G, r) - RW(A_S, r);
        RW(dsarr, r) = gd + c;
      }
    };
    auto ftb_primal = [&](const double* du_, int dsarr) {
      double a = 1.0;
      if (hasu) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (is_lo(xL[i]) && du_[i] < 0.0) a = fmin(a, -tau * (u[i] - xL[i]) / du_[i]);
          if (is_hi(xU[i]) && du_[i] > 0.0) a = fmin(a, tau * (xU[i] - u[i]) / du_[i]);
        }
      }
      if (act) for (int r = 0; r < R; ++r) {
        const double sv = RW(A_S, r), dsv = RW(dsarr, r), lo = RW(A_DL, r), hi = RW(A_DU, r);
        if (is_lo(lo) && dsv < 0.0) a = fmin(a, -tau * (sv - lo) / dsv);
        if (is_hi(hi) && dsv > 0.0) a = fmin(a, tau * (hi - sv) / dsv);
      }
      return warp_min(a);
    };
    // dual steps are recomputed from (du, ds) where needed:  dz = (mu -+ z*d)/slack - z
    auto ftb_dual = [&](const double* du_, int dsarr) {
      double a = 1.0;
      if (hasu) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (is_lo(xL[i])) { const double dz = (mu - zL[i] * du_[i]) / (u[i] - xL[i]) - zL[i]; if (dz < 0.0) a = fmin(a, -tau * zL[i] / dz); }
          if (is_hi(xU[i])) { const double dz = (mu + zU[i] * du_[i]) / (xU[i] - u[i]) - zU[i]; if (dz < 0.0) a = fmin(a, -tau * zU[i] / dz); }
        }
      }
      if (act) for (int r = 0; r < R; ++r) {
        const double sv = RW(A_S, r), dsv = RW(dsarr, r), lo = RW(A_DL, r), hi = RW(A_DU, r);
        if (is_lo(lo)) { const double v = RW(A_VL, r), dz = (mu - v * dsv) / (sv - lo) - v; if (dz < 0.0) a = fmin(a, -tau * v / dz); }
        if (is_hi(hi)) { const double v = RW(A_VU, r), dz = (mu + v * dsv) / (hi - sv) - v; if (dz < 0.0) a = fmin(a, -tau * v / dz); }
      }
      return warp_min(a);
    };
    row_steps(mydx, A_DS, false);
    const double a_pr_max = ftb_primal(mydu, A_DS);
    double a_du = ftb_dual(mydu, A_DS);
    // ---- line-search reference quantities
    double theta = 0.0, gbd = 0.0, tiny_l = 0.0;
    if (hasu) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const bool hl = is_lo(xL[i]), hu = is_hi(xU[i]);
        double bgv = (hl ? -mu / (u[i] - xL[i]) : 0.0) + (hu ? mu / (xU[i] - u[i]) : 0.0);
        if (hl && !hu) bgv += o.kappa_d * mu; if (hu && !hl) bgv -= o.kappa_d * mu;
        gbd += bgv * mydu[i];
        tiny_l = fmax(tiny_l, fabs(mydu[i]) / (1.0 + fabs(u[i])));
      }
#pragma unroll
      for (int v = 0; v < 6; ++v) gbd += gl[v] * mydx[cost_state(v)];
    }
    if (act) for (int r = 0; r < R; ++r) {
      const double sv = RW(A_S, r), lo = RW(A_DL, r), hi = RW(A_DU, r), dsv = RW(A_DS, r);
      const bool hl = is_lo(lo), hu = is_hi(hi);
      double bgv = (hl ? -mu / (sv - lo) : 0.0) + (hu ? mu / (hi - sv) : 0.0);
      if (hl && !hu) bgv += o.kappa_d * mu; if (hu && !hl) bgv -= o.kappa_d * mu;
      gbd += bgv * dsv;
      theta += fabs(RW(A_G, r) - sv);
      tiny_l = fmax(tiny_l, fabs(dsv) / (1.0 + fabs(sv)));
    }
    theta = warp_sum(theta); gbd = warp_sum(gbd); tiny_l = warp_max(tiny_l);
    const double phi = f - mu * LB + o.kappa_d * mu * DT;
    bool tiny = tiny_l <= o.tiny_step_tol && theta <= 1e-4;

    // trial point evaluation: u + alpha*du_, s + alpha*ds[dsarr]
    Stage stt; double ut[6], f_t, th_t, LB_t, DT_t;
    auto eval_trial = [&](double alpha, const double* du_, int dsarr) {
      ++n_ls;
#pragma unroll
      for (int i = 0; i < 6; ++i) ut[i] = u[i] + alpha * du_[i];
      rollout(pr, X0, ut, lane, stt);
      f_t = cost_sum(stt);
      eval_g(stt, A_GT);
      double th = 0.0;
      if (act) for (int r = 0; r < R; ++r) th += fabs(RW(A_GT, r) - (RW(A_S, r) + alpha * RW(dsarr, r)));
      th_t = warp_sum(th);
      barrier_parts(ut, 1, alpha, dsarr, LB_t, DT_t);
    };
    auto is_ftype = [&](double a) { return gbd < 0.0 && a * pow(-gbd, o.s_phi) > o.delta * pow(theta, o.s_theta); };
    auto armijo = [&](double a, double ph_t) { return cmp_le(ph_t - phi, o.eta_phi * a * gbd, phi); };
    auto acceptable = [&](double a_test, double th_, double ph_) {
      if (!isfinite(th_) || !isfinite(ph_)) return false;
      if (th_ > theta_max) return false;
      bool acc;
      if (a_test > 0.0 && is_ftype(a_test) && theta <= theta_min) acc = armijo(a_test, ph_);
      else acc = cmp_le(th_, (1.0 - o.gamma_theta) * theta, theta) || cmp_le(ph_ - phi, -o.gamma_phi * theta, phi);
      if (!acc) return false;
      for (int e = 0; e < nfilt; ++e) if (!(th_ < filt[2 * e] || ph_ < filt[2 * e + 1])) return false;
      return true;
    };
    double alpha = a_pr_max, alpha_test = a_pr_max; bool accepted = false, used_soc = false;
    double phi_t = 0.0;
    double dxs[8], dus[6];   // SOC direction
    if (tiny) {
      eval_trial(alpha, mydu, A_DS); accepted = true; ++tiny_count; tiny_flag = true;
      phi_t = f_t - mu * LB_t + o.kappa_d * mu * DT_t;
      if (tiny_count >= 2 && mu <= mu_floor) { status = NMPC_STEP_TOO_SMALL; break; }
    } else {
      tiny_count = 0;
      double amin = o.gamma_theta;
      if (gbd < 0.0) {
        amin = fmin(o.gamma_theta, o.gamma_phi * theta / (-gbd));
        if (theta <= theta_min) amin = fmin(amin, o.delta * pow(theta, o.s_theta) / pow(-gbd, o.s_phi));
      }
      amin *= o.alpha_min_frac;
      bool first = true;
      while (alpha > amin || first) {
        eval_trial(alpha, mydu, A_DS);
        phi_t = f_t - mu * LB_t + o.kappa_d * mu * DT_t;
        alpha_test = alpha;
        if (acceptable(alpha, th_t, phi_t)) { accepted = true; break; }
        if (first && o.max_soc > 0 && th_t >= theta && isfinite(th_t)) {
          // ---- second-order correction
          double a_soc = alpha, th_prev = th_t;
          if (act) for (int r = 0; r < R; ++r) RW(A_CSOC, r) = RW(A_G, r) - RW(A_S, r);
          int dsprev = A_DS;
          for (int kk = 0; kk < o.max_soc; ++kk) {
            // c_soc = a_soc * c_soc + c(trial); q' = grad l + G^T (D_s c_soc + barrier gradient)
            if (act) {
              double q[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) q[i] = 0.0;
#pragma unroll
              for (int v = 0; v < 6; ++v) q[cost_state(v)] = gl[v];
              for (int r = 0; r < R; ++r) {
                const double sv = RW(A_S, r), lo = RW(A_DL, r), hi = RW(A_DU, r), dc = RW(A_DC, r);
                const bool hl = is_lo(lo), hu = is_hi(hi);
                const double cs = a_soc * RW(A_CSOC, r) + (RW(A_GT, r) - (sv + a_soc * RW(dsprev, r)));
                RW(A_CSOC, r) = cs;
                const double sig = (hl ? RW(A_VL, r) / (sv - lo) : 0.0) + (hu ? RW(A_VU, r) / (hi - sv) : 0.0) + dw;
                double bg = (hl ? -mu / (sv - lo) : 0.0) + (hu ? mu / (hi - sv) : 0.0);
                if (hl && !hu) bg += o.kappa_d * mu; if (hu && !hl) bg -= o.kappa_d * mu;
                const double yh = sig * cs + bg;
                if (r < 5) q[box_state(r)] += dc * yh;
                else {
                  const int jn = r - 5;
                  const double ddx = st.X[0] - obs[3 * jn], ddy = st.X[1] - obs[3 * jn + 1];
                  const double iD = 1.0 / sqrt(ddx * ddx + ddy * ddy);
                  q[0] += -dc * yh * ddx * iD; q[1] += -dc * yh * ddy * iD;
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) LQ(LQ_Q2 + i) = q[i];
            }
            __syncwarp();
            riccati_resolve(pr, lq, ric, lane);
            riccati_forward(pr, lq, ric, true, lane, dxs, dus);
            row_steps(dxs, A_DS2, true);
            a_soc = ftb_primal(dus, A_DS2);
            eval_trial(a_soc, dus, A_DS2);
            dsprev = A_DS2;
            const double ph_s = f_t - mu * LB_t + o.kappa_d * mu * DT_t;
            if (acceptable(alpha, th_t, ph_s)) { accepted = true; used_soc = true; phi_t = ph_s; alpha = a_soc; ++n_soc; break; }
            if (!(th_t <= o.kappa_soc * th_prev)) break;
            th_prev = th_t;
          }
          if (accepted) break;
        }
        first = false;
        alpha *= o.alpha_red;
      }
    }
    if (!accepted) { status = NMPC_RESTORATION_NEEDED; break; }
    // ---- filter augmentation
    if (!tiny && !(is_ftype(alpha_test) && armijo(alpha_test, phi_t))) {
      if (lane == 0) {
        if (nfilt == FILT_CAP) for (int e = 0; e < 2 * (FILT_CAP - 1); ++e) filt[e] = filt[e + 2];   // drop the oldest entry
        const int at = nfilt == FILT_CAP ? FILT_CAP - 1 : nfilt;
        filt[2 * at] = (1.0 - o.gamma_theta) * theta; filt[2 * at + 1] = phi - o.gamma_phi * theta;
      }
      if (nfilt < FILT_CAP) ++nfilt;
      __syncwarp();
    }
    if (A.dbg && lane == 0 && iter < A.dbg_rows) {
      double* L = A.dbg + ((size_t)b * A.dbg_rows + iter) * 8;
      L[0] = mu; L[1] = f / df; L[2] = pr_inf; L[3] = du_inf; L[4] = dw; L[5] = alpha; L[6] = a_du; L[7] = (double)(n_ls - ls_before);
    }
    // ---- accept the trial point
    const double* du_acc = used_soc ? dus : mydu;
    const int ds_acc = used_soc ? A_DS2 : A_DS;
    if (used_soc) a_du = ftb_dual(dus, A_DS2);
    if (hasu) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const bool hl = is_lo(xL[i]), hu = is_hi(xU[i]);
        const double sl = u[i] - xL[i], su = xU[i] - u[i];
        double zl = zL[i], zu = zU[i];
        if (hl) zl += a_du * ((mu - zl * du_acc[i]) / sl - zl);
        if (hu) zu += a_du * ((mu + zu * du_acc[i]) / su - zu);
        u[i] = ut[i];
        if (hl) { const double s2 = u[i] - xL[i]; zl = fmax(fmin(zl, o.kappa_sigma * mu / s2), mu / (o.kappa_sigma * s2)); }
        if (hu) { const double s2 = xU[i] - u[i]; zu = fmax(fmin(zu, o.kappa_sigma * mu / s2), mu / (o.kappa_sigma * s2)); }
        zL[i] = zl; zU[i] = zu;
      }
    }
    if (act) for (int r = 0; r < R; ++r) {
      const double sv = RW(A_S, r), lo = RW(A_DL, r), hi = RW(A_DU, r), dsv = RW(ds_acc, r);
      const bool hl = is_lo(lo), hu = is_hi(hi);
      double vl = RW(A_VL, r), vu = RW(A_VU, r);
      const double sig = (hl ? vl / (sv - lo) : 0.0) + (hu ? vu / (hi - sv) : 0.0) + dw;
      double bg = (hl ? -mu / (sv - lo) : 0.0) + (hu ? mu / (hi - sv) : 0.0);
      if (hl && !hu) bg += o.kappa_d * mu; if (hu && !hl) bg -= o.kappa_d * mu;
      const double yv = RW(A_Y, r);
      const double dy = sig * dsv + (-yv + bg);
      RW(A_Y, r) = yv + alpha * dy;
      if (hl) vl += a_du * ((mu - vl * dsv) / (sv - lo) - vl);
      if (hu) vu += a_du * ((mu + vu * dsv) / (hi - sv) - vu);
      const double sn = sv + alpha * dsv;
      if (hl) { const double s2 = sn - lo; vl = fmax(fmin(vl, o.kappa_sigma * mu / s2), mu / (o.kappa_sigma * s2)); }
      if (hu) { const double s2 = hi - sn; vu = fmax(fmin(vu, o.kappa_sigma * mu / s2), mu / (o.kappa_sigma * s2)); }
      RW(A_S, r) = sn; RW(A_VL, r) = vl; RW(A_VU, r) = vu; RW(A_G, r) = RW(A_GT, r);
    }
    st = stt; f = f_t; LB = LB_t; DT = DT_t;
    ++iter;
    __syncwarp();
  }

  // ---------------- outputs: honour original bounds, unscale ----------------------------------
  const int nw = NU * N, ng = R * S;
  if (hasu) {
#pragma unroll
    for (int i = 0; i < 6; ++i) u[i] = fmin(fmax(u[i], A.lbx[NU * lane + i]), A.ubx[NU * lane + i]);
    double* xo = A.x + (size_t)b * nw + NU * lane;
#pragma unroll
    for (int i = 0; i < 6; ++i) xo[i] = u[i];
    if (A.lam_x) {
      double* lo = A.lam_x + (size_t)b * nw + NU * lane;
#pragma unroll
      for (int i = 0; i < 6; ++i) lo[i] = (zU[i] - zL[i]) / df;
    }
  }
  rollout(pr, X0, u, lane, st);
  {
    const double l = hasu ? stage_cost(pr, st.X, xt, yt) : 0.0;
    const double fu = warp_sum(l);
    if (lane == 0) {
      if (A.f) A.f[b] = fu;
      if (A.status) A.status[b] = status;
      if (A.iters) A.iters[b] = iter;
      if (A.stats) { atomicAdd(&A.stats[0], n_fact); atomicAdd(&A.stats[1], n_ls); atomicAdd(&A.stats[2], n_soc); }
    }
  }
  if (act) {
    if (A.g) {
      double* go = A.g + (size_t)b * ng + lane * R;
#pragma unroll
      for (int r = 0; r < 5; ++r) go[r] = st.X[box_state(r)];
      for (int jn = 0; jn < n_obs; ++jn) {
        const double dx_ = st.X[0] - obs[3 * jn], dy_ = st.X[1] - obs[3 * jn + 1];
        go[5 + jn] = obs[3 * jn + 2] - sqrt(dx_ * dx_ + dy_ * dy_);
      }
    }
    if (A.lam_g) {
      double* lo = A.lam_g + (size_t)b * ng + lane * R;
      for (int r = 0; r < R; ++r) lo[r] = RW(A_Y, r) * RW(A_DC, r) / df;
    }
  }
  __syncwarp();
#undef RW
#undef LQ
}

}  // namespace nmpc
