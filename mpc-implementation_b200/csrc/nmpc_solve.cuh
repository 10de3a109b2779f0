// nmpc_solve.cuh -- one-warp-per-instance primal-dual interior-point solve (the hot path).
//
// Replaces  sol = solver(x0,lbx,ubx,lbg,ubg,p)  (Python/NMPC_TT.py:358-365; CasADi -> IPOPT -> MUMPS) for the
// reference's UAV target-tracking NLP.  Algorithm = IPOPT's (Waechter & Biegler 2006) with the scripts'
// options (NMPC_TT.py:257-265): slack form g(w) - s = 0, relaxed bounds, gradient-based scaling,
// monotone barrier, fraction-to-boundary, inertia correction, filter line search + second-order
// correction, scaled optimality-error termination.  The whole iteration loop runs inside the kernel.
//
// Structure: the iterate lives in the warp's slice of shared memory at compile-time offsets (Lay<N, NOBS>); every
// phase is a __noinline__ function template that loads its operands, works in registers and stores back, so each
// heavy code sequence exists exactly once and every workspace access is an LDS/STS with an immediate offset.  A block
// holds Lay::WPB warps (8 at N = 15, n_obs = 3) that start every IPM iteration together (align_warps) so that they
// share instruction-cache lines.  One launch is one whole closed-loop step when asked (nmpc_solve_and_step: the shift
// runs in ph_output) and also prepares the next call on the handle (queue / counters, longest-first fetch order).
#pragma once
#include "nmpc_device.cuh"
#include "nmpc_riccati.cuh"

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
  const double* weights;             // optional per-instance cost weights [B][2] = (w1, w2); NULL = spec weights
  const double* tgt;                 // optional per-instance, per-stage predicted target [B][N][2]; NULL = p[8:10]
  // fused closed-loop shift (nmpc_solve_and_step): when step_p != NULL the warp that solved instance b also applies
  // shift_timestep to it -- p[b] and u_warm[b] in place, FOV centre, error term -- and no nmpc_step launch is needed
  double *step_p, *step_u; const double* step_vw; double *step_fov, *step_err;
  int32_t* iters_keep;               // handle-owned copy of iters[] (drives the next call's fetch order)
  const int32_t* order;              // optional processing order (longest-first scheduling); NULL = 0..B-1
  int* counter;                      // work queue of THIS call
  // Bookkeeping for the NEXT call on the same handle is done by this launch, so that a solve is ONE launch: the first
  // thread resets the next call's queue / done / work counters (calls on a handle are stream-ordered), and the warp that
  // finishes the last instance writes the next call's longest-first fetch order from this call's iteration counts.
  int *counter_next, *done, *done_next;
  unsigned long long* stats_next;
  int32_t* order_out;                // may be NULL (no automatic ordering)
  unsigned long long* stats;         // [3]: factorizations, ls trials, soc accepted
  double* ric; int ric_stride;       // L2-resident Riccati scratch, one slice per resident warp
  const unsigned* ricmap;            // per-lane ownership maps of the factorisation (nmpc_riccati.cuh: ric_map_build)
  double* dbg; int dbg_rows;         // optional per-iteration log [B][dbg_rows][8] (tests only)
  int align_group;                   // warps that start each IPM iteration together (0 = no alignment, else divides WPB)
};

constexpr double EPSM = 2.220446049250313e-16;
__device__ __forceinline__ bool cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * EPSM * fabs(bas); }

__device__ __forceinline__ double push_in(double v, double lo, double hi, bool hl, bool hu, double k1, double k2) {
  if (hl && hu) {
    const double pl = fmin(k1 * fmax(1.0, fabs(lo)), k2 * (hi - lo));
    const double pu = fmin(k1 * fmax(1.0, fabs(hi)), k2 * (hi - lo));
    v = fmax(v, lo + pl); v = fmin(v, hi - pu);
  } else if (hl) v = fmax(v, lo + k1 * fmax(1.0, fabs(lo)));
  else if (hu) v = fmin(v, hi - k1 * fmax(1.0, fabs(hi)));
  return v;
}

struct Bnd { double lo, hi; bool hl, hu; };
// Relaxed bounds of control i of stage k (NMPC_TT.py:294-306) and of the scaled row r of stage k (:275-291).
// NOTE: these are inlined into several phase functions; the explicit round-to-nearest intrinsics keep the compiler
// from contracting the expressions into FMAs differently per copy.  A bound that differs by one ulp between the
// function that builds the Newton system and the one that updates the multipliers is a 1e-6 RELATIVE error of a
// 1e-10 slack, i.e. a 1e-5 error in a multiplier -- enough to stall the end game (seen with v2).
__device__ __forceinline__ double relaxed_lo(double lo, double relax) { return lo > -1e19 ? __dsub_rn(lo, __dmul_rn(relax, fmax(1.0, fabs(lo)))) : -CUDART_INF; }
__device__ __forceinline__ double relaxed_hi(double hi, double relax) { return hi < 1e19 ? __dadd_rn(hi, __dmul_rn(relax, fmax(1.0, fabs(hi)))) : CUDART_INF; }
__device__ __forceinline__ Bnd ctl_bounds(const SolveArgs& A, int k, int i) {
  Bnd b; b.lo = relaxed_lo(__ldg(A.lbx + NU * k + i), A.o.bound_relax); b.hi = relaxed_hi(__ldg(A.ubx + NU * k + i), A.o.bound_relax);
  b.hl = b.lo > -CUDART_INF; b.hu = b.hi < CUDART_INF;
  return b;
}
template <class L>
__device__ __forceinline__ Bnd row_bounds(const SolveArgs& A, int k, int r, double dc) {
  const double lo = __ldg(A.lbg + k * L::R + r), hi = __ldg(A.ubg + k * L::R + r);
  Bnd b;
  b.lo = relaxed_lo(lo > -1e19 ? __dmul_rn(dc, lo) : lo, A.o.bound_relax);
  b.hi = relaxed_hi(hi < 1e19 ? __dmul_rn(dc, hi) : hi, A.o.bound_relax);
  b.hl = b.lo > -CUDART_INF; b.hu = b.hi < CUDART_INF;
  return b;
}

// Alignment barrier over the group of `g` consecutive warps this warp belongs to (named barrier 1 + group index),
// OR-reducing `working` over the group.  Every warp of the group calls it the same number of times: working warps
// once per IPM iteration, warps that ran out of work in a loop until the reduction says nobody is working -- the
// exit decision comes out of the barrier itself, so no warp can leave while another still waits.  g == 0: no alignment.
__device__ __forceinline__ int align_warps(int g, int working) {
  if (g == 0) return 0;
  unsigned out;
  const unsigned id = 1 + (threadIdx.x >> 5) / g, nt = 32 * g;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
               : "=r"(out) : "r"((unsigned)working), "r"(id), "r"(nt) : "memory");
  return (int)out;
}

#define LV(e) smem[L::LV0 + (e) * L::S + lane]
#define RW(arr, r) smem[L::RW0 + ((arr) * L::R + (r)) * L::S + lane]
#define LQ(e) smem[L::LQ0 + (e) * L::S + lane]
#define SOC(e) smem[L::soc(e) + lane]
#define RES(i) smem[L::RES0 + (i)]
#define PAR(i) smem[L::PAR0 + (i)]
#define SOC_DS2 0
#define SOC_CSOC (L::R)
#define SOC_CT (2 * L::R)
#define SOC_DUS (3 * L::R)
#define SOC_Q2 (3 * L::R + 6)

// T-scaled non-zeros of the dynamics Jacobian A_k - I of this lane's stage
struct Dyn { double e03, e13, e23, e04, e14; };
__device__ __forceinline__ Dyn dyn_entries(const Stage& s, double tv) {
  Dyn d; d.e03 = -tv * s.cps * s.sth; d.e13 = -tv * s.sps * s.sth; d.e23 = tv * s.cth;
  d.e04 = -tv * s.sps * s.cth; d.e14 = tv * s.cps * s.cth;
  return d;
}
// adjoint recursion lam_k = a_k + A_k^T lam_{k+1} by suffix scans; returns lam_{k+1} in lamn
__device__ __forceinline__ void adjoint(const double* a, const Dyn& d, bool act, int lane, double* lamn) {
  double v[6] = {act ? a[0] : 0.0, act ? a[1] : 0.0, act ? a[2] : 0.0, act ? a[5] : 0.0, act ? a[6] : 0.0, act ? a[7] : 0.0};
  rscan_incl<6>(v, lane);
  lamn[0] = shfl_next(v[0], lane); lamn[1] = shfl_next(v[1], lane); lamn[2] = shfl_next(v[2], lane);
  lamn[5] = shfl_next(v[3], lane); lamn[6] = shfl_next(v[4], lane); lamn[7] = shfl_next(v[5], lane);
  double w[2] = {(act ? a[3] : 0.0) + d.e03 * lamn[0] + d.e13 * lamn[1] + d.e23 * lamn[2],
                 (act ? a[4] : 0.0) + d.e04 * lamn[0] + d.e14 * lamn[1]};
  rscan_incl<2>(w, lane);
  lamn[3] = shfl_next(w[0], lane); lamn[4] = shfl_next(w[1], lane);
}

// obstacle row jn of this lane's stage: value (r_u + r_o) - ||(x,y) - c_j||, unit normal, 1/distance   NMPC_TT.py:241-243
template <class L>
__device__ __forceinline__ double obs_value(const double* X, int jn, double& nx, double& ny, double& iD) {
  const double dx_ = X[0] - smem[L::OBS0 + 3 * jn], dy_ = X[1] - smem[L::OBS0 + 3 * jn + 1];
  const double d2 = __fma_rn(dx_, dx_, __dmul_rn(dy_, dy_));
  const double D = sqrt(d2);
  iD = rcp(D); nx = dx_ * iD; ny = dy_ * iD;
  return smem[L::OBS0 + 3 * jn + 2] - D;
}
// Visit the rows of this lane's stage: the five box rows fully unrolled (static state index si), the obstacle
// rows in a rolled loop.  body(r, box, si, gu, nx, ny, iD) with gu the unscaled row value.
#define FOR_ROWS(X, body)                                                                           \
  do {                                                                                              \
    _Pragma("unroll") for (int r_ = 0; r_ < 5; ++r_) {                                              \
      const int si_ = r_ == 0 ? 2 : (r_ == 1 ? 3 : r_ + 3);                                         \
      body(r_, true, si_, (X)[si_], 0.0, 0.0, 0.0);                                                 \
    }                                                                                               \
    _Pragma("unroll 1") for (int jn_ = 0; jn_ < L::NOBS; ++jn_) {                                   \
      double nx_, ny_, iD_;                                                                         \
      const double gu_ = obs_value<L>((X), jn_, nx_, ny_, iD_);                                     \
      body(5 + jn_, false, 0, gu_, nx_, ny_, iD_);                                                  \
    }                                                                                               \
  } while (0)

// target position seen by this lane's stage: the reference keeps (x_t, y_t) = p[8:10] over the horizon
// (NMPC_TT.py:219-220); with a predicted trajectory (nmpc_set_target_trajectory) stage k has its own
template <class L>
__device__ __forceinline__ double2 stage_target(const SolveArgs& A, int lane) {
  if (!A.tgt || lane >= L::N) return make_double2(PAR(8), PAR(9));
  const double* t = A.tgt + ((size_t)PAR(NPAR + 2) * L::N + lane) * 2;
  return make_double2(__ldg(t), __ldg(t + 1));
}

// ---------------------------------------------------------------------------------------------------
// load one instance: p -> PAR, warm start -> LV_U, obstacle table, unit scaling
template <class L>
__device__ __noinline__ void ph_load(const SolveArgs& A, int b, int lane) {
  if (lane < NPAR) PAR(lane) = A.p[(size_t)b * NPAR + lane];
  if (lane == NPAR) PAR(NPAR) = A.weights ? A.weights[2 * (size_t)b] : A.pr.w1;
  if (lane == NPAR + 1) PAR(NPAR + 1) = A.weights ? A.weights[2 * (size_t)b + 1] : A.pr.w2;
  if (lane == NPAR + 2) PAR(NPAR + 2) = (double)b;      // instance index, for the per-stage target lookup
  const double* ob = A.obs + (A.obs_per_instance ? (size_t)b * 3 * L::NOBS : 0);
  for (int i = lane; i < 3 * L::NOBS; i += 32) smem[L::OBS0 + i] = ob[i];
  if (lane <= L::N) {
    const double* xx = A.x0 + (size_t)b * (NU * L::N) + NU * lane;
#pragma unroll
    for (int i = 0; i < 6; ++i) { LV(LV_U + i) = lane < L::N ? xx[i] : 0.0; LV(LV_DU + i) = 0.0; }
#pragma unroll 1
    for (int r = 0; r < L::R; ++r) { RW(A_DC, r) = 1.0; RW(A_DS, r) = 0.0; }
  }
  __syncwarp();
}

// gradient-based scaling at the user's starting point (IPOPT nlp_scaling_method = gradient-based): returns df, sets DC
template <class L>
__device__ __noinline__ double ph_scaling(const SolveArgs& A, int lane) {
  const Prob& pr = A.pr; constexpr int N = L::N; const double T = pr.T;
  const bool act = lane <= N, hasu = lane < N;
  double u[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) u[i] = act ? LV(LV_U + i) : 0.0;
  Stage st; rollout(pr, &PAR(0), u, lane, st);
  const Dyn dy = dyn_entries(st, hasu ? T * u[0] : 0.0);
  double gl[6], Hl[21], a[8], lamn[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.0;
  if (hasu && lane >= 1) {
    stage_cost_d2(with_weights(pr, PAR(NPAR), PAR(NPAR + 1)), st.X, stage_target<L>(A, lane).x, stage_target<L>(A, lane).y, gl, Hl);
#pragma unroll
    for (int v = 0; v < 6; ++v) a[cost_state(v)] = gl[v];
  }
  adjoint(a, dy, act, lane, lamn);
  double gmax = 0.0;
  if (hasu) {
    gmax = fabs(T * (st.cps * st.cth * lamn[0] + st.sps * st.cth * lamn[1] + st.sth * lamn[2]));
#pragma unroll
    for (int r = 1; r < 6; ++r) gmax = fmax(gmax, fabs(T * lamn[r + 2]));
  }
  gmax = warp_max(gmax);
  // row maxima of the Jacobian: |d g_{k,i} / d u_j| for j < k
  const double d0 = st.cps * st.cth, d1 = st.sps * st.cth, d2 = st.sth;
  double c3[3] = {dy.e03, dy.e13, dy.e23}, c4[2] = {dy.e04, dy.e14};
  scan_excl<3>(c3, lane); scan_excl<2>(c4, lane);
  double zmax = 0.0;
  // scratch in row arrays that are not live yet: A_G = row max, A_IL / A_IU = obstacle normal
  if (act) {
#pragma unroll 1
    for (int jn = 0; jn < L::NOBS; ++jn) {
      double nx, ny, iD; obs_value<L>(st.X, jn, nx, ny, iD);
      RW(A_G, 5 + jn) = 0.0; RW(A_IL, 5 + jn) = nx; RW(A_IU, 5 + jn) = ny;
    }
  }
#pragma unroll 1
  for (int jj = 0; jj < N; ++jj) {
    const double dj0 = __shfl_sync(FULL, d0, jj), dj1 = __shfl_sync(FULL, d1, jj), dj2 = __shfl_sync(FULL, d2, jj);
    const double a30 = __shfl_sync(FULL, c3[0], jj + 1), a31 = __shfl_sync(FULL, c3[1], jj + 1), a32 = __shfl_sync(FULL, c3[2], jj + 1);
    const double a40 = __shfl_sync(FULL, c4[0], jj + 1), a41 = __shfl_sync(FULL, c4[1], jj + 1);
    if (act && jj < lane) {
      const double pv0 = T * dj0, pv1 = T * dj1, pv2 = T * dj2;
      const double pt0 = T * (c3[0] - a30), pt1 = T * (c3[1] - a31), pt2 = T * (c3[2] - a32);
      const double pp0 = T * (c4[0] - a40), pp1 = T * (c4[1] - a41);
      zmax = fmax(zmax, fmax(fabs(pv2), fabs(pt2)));
#pragma unroll 1
      for (int jn = 0; jn < L::NOBS; ++jn) {
        const double nx = RW(A_IL, 5 + jn), ny = RW(A_IU, 5 + jn);
        const double m1 = fabs(nx * pv0 + ny * pv1), m2 = fabs(nx * pt0 + ny * pt1), m3 = fabs(nx * pp0 + ny * pp1);
        RW(A_G, 5 + jn) = fmax(RW(A_G, 5 + jn), fmax(m1, fmax(m2, m3)));
      }
    }
  }
  if (act) {
    const double mg = A.o.max_grad, smin = A.o.scal_min;
    RW(A_DC, 0) = zmax > mg ? fmax(smin, mg / zmax) : 1.0;
    const double lin = lane >= 1 ? T : 0.0, sl = lin > mg ? fmax(smin, mg / lin) : 1.0;
#pragma unroll
    for (int r = 1; r < 5; ++r) RW(A_DC, r) = sl;
#pragma unroll 1
    for (int jn = 0; jn < L::NOBS; ++jn) { const double m = RW(A_G, 5 + jn); RW(A_DC, 5 + jn) = m > mg ? fmax(smin, mg / m) : 1.0; }
  }
  __syncwarp();
  return gmax > A.o.max_grad ? fmax(A.o.scal_min, A.o.max_grad / gmax) : 1.0;
}

// starting point: push controls and slacks inside their bounds, unit bound multipliers; returns #finite bounds
template <class L>
__device__ __noinline__ int ph_start(const SolveArgs& A, int lane) {
  const Prob& pr = A.pr; constexpr int N = L::N;
  const bool act = lane <= N, hasu = lane < N;
  double u[6]; int nz = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    u[i] = act ? LV(LV_U + i) : 0.0;
    if (hasu) {
      const Bnd b = ctl_bounds(A, lane, i);
      u[i] = push_in(u[i], b.lo, b.hi, b.hl, b.hu, A.o.bound_push, A.o.bound_frac);
      LV(LV_U + i) = u[i]; LV(LV_ZL + i) = b.hl ? 1.0 : 0.0; LV(LV_ZU + i) = b.hu ? 1.0 : 0.0;
      nz += (b.hl ? 1 : 0) + (b.hu ? 1 : 0);
    } else if (act) { LV(LV_ZL + i) = 0.0; LV(LV_ZU + i) = 0.0; }
  }
  Stage st; rollout(pr, &PAR(0), u, lane, st);
  if (act) {
    auto body = [&](const int r, const bool box, const int si, const double gu, double nx, double ny, double iD) {
      const double dc = RW(A_DC, r);
      const double g = __dmul_rn(dc, gu);
      const Bnd b = row_bounds<L>(A, lane, r, dc);
      const double s = push_in(g, b.lo, b.hi, b.hl, b.hu, A.o.bound_push, A.o.bound_frac);
      RW(A_G, r) = g; RW(A_S, r) = s; RW(A_Y, r) = 0.0;
      RW(A_VL, r) = b.hl ? 1.0 : 0.0; RW(A_VU, r) = b.hu ? 1.0 : 0.0;
      RW(A_IL, r) = b.hl ? rcp(s - b.lo) : 0.0; RW(A_IU, r) = b.hu ? rcp(b.hi - s) : 0.0;
      nz += (b.hl ? 1 : 0) + (b.hu ? 1 : 0);
    };
    FOR_ROWS(st.X, body);
  }
  __syncwarp();
  return __reduce_add_sync(FULL, nz);
}

// ---------------------------------------------------------------------------------------------------
// derivatives at the current point: stage Hessian blocks / gradients of the LQ sub-problem into LQ, optimality
// error ingredients into RES.  ls = least-squares multiplier system (W = 0, Sigma = I, rhs = gradient of L).
template <class L>
__device__ __noinline__ void ph_derivs(const SolveArgs& A, int lane, bool ls, double df) {
  const Prob& pr = A.pr; constexpr int N = L::N; const double T = pr.T;
  const bool act = lane <= N, hasu = lane < N;
  double u[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) u[i] = act ? LV(LV_U + i) : 0.0;
  Stage st; rollout(pr, &PAR(0), u, lane, st);
  const Dyn dy = dyn_entries(st, hasu ? T * u[0] : 0.0);
  double gl[6], Hl[21];
#pragma unroll
  for (int v = 0; v < 6; ++v) gl[v] = 0.0;
#pragma unroll
  for (int e = 0; e < 21; ++e) Hl[e] = 0.0;
  double l = 0.0;
  if (hasu) l = stage_cost_d2(with_weights(pr, PAR(NPAR), PAR(NPAR + 1)), st.X, stage_target<L>(A, lane).x, stage_target<L>(A, lane).y, gl, Hl);
  if (lane == 0) {   // stage 0 is constant in w
#pragma unroll
    for (int v = 0; v < 6; ++v) gl[v] = 0.0;
#pragma unroll
    for (int e = 0; e < 21; ++e) Hl[e] = 0.0;
  }
  const double fsum = df * warp_sum(l);
  double a[8], qa[8], qb[8], qd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 0.0; qa[i] = 0.0; qb[i] = 0.0; qd[i] = 0.0; }
  double du_l = 0.0, pr_l = 0.0, sumy = 0.0, sumz = 0.0, viol = 0.0, pmax = 0.0, pmin = CUDART_INF;
  if (act) {
#pragma unroll
    for (int i = 0; i < 8; ++i) LV(LV_X + i) = st.X[i];
#pragma unroll
    for (int v = 0; v < 6; ++v) { gl[v] *= df; LV(LV_GL + v) = gl[v]; a[cost_state(v)] = gl[v]; qa[cost_state(v)] = gl[v]; }
    double q66[21], q22 = 0.0, nn[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int e = 0; e < 21; ++e) q66[e] = ls ? 0.0 : df * Hl[e];
    const double kd = A.o.kappa_d;
    auto body = [&](const int r, const bool box, const int si, const double gu, double nx, double ny, double iD) {
      const double dc = RW(A_DC, r);
      const double g = __dmul_rn(dc, gu);
      const double s = RW(A_S, r), y = RW(A_Y, r), vl = RW(A_VL, r), vu = RW(A_VU, r), il = RW(A_IL, r), iu = RW(A_IU, r);
      const bool hl = il > 0.0, hu = iu > 0.0;
      RW(A_G, r) = g;
      const double c = g - s;
      const double sig = ls ? 1.0 : vl * il + vu * iu;
      double beta = iu - il;                              // barrier gradient per unit mu (with damping)
      if (hl && !hu) beta += kd;
      if (hu && !hl) beta -= kd;
      const double ya = ls ? (vu - vl) : sig * c;         // delta_w- and mu-free part of y-hat
      const double w = dc * dc * sig;
      if (box) {
        a[si] += dc * y; qa[si] += dc * ya; qb[si] += dc * beta; qd[si] += dc * c;
        if (si == 3) q22 += w; else q66[tri(si < 3 ? si : si - 2, si < 3 ? si : si - 2)] += w;
        LQ(LQ_DG + r) = dc * dc;
      } else {
        const double cur = ls ? 0.0 : -y * dc * iD;        // y * d2h,  d2h = -(I - n n^T)/D
        q66[0] += w * nx * nx + cur * (1.0 - nx * nx);
        q66[1] += w * nx * ny - cur * nx * ny;
        q66[2] += w * ny * ny + cur * (1.0 - ny * ny);
        nn[0] += dc * dc * nx * nx; nn[1] += dc * dc * nx * ny; nn[2] += dc * dc * ny * ny;
        const double gy = -dc * y, ga = -dc * ya, gb = -dc * beta, gd = -dc * c;
        a[0] += gy * nx; a[1] += gy * ny; qa[0] += ga * nx; qa[1] += ga * ny;
        qb[0] += gb * nx; qb[1] += gb * ny; qd[0] += gd * nx; qd[1] += gd * ny;
      }
      // optimality-error ingredients
      du_l = fmax(du_l, fabs(-y - vl + vu)); pr_l = fmax(pr_l, fabs(c)); sumy += fabs(y); sumz += vl + vu;
      const Bnd b = row_bounds<L>(A, lane, r, dc);
      if (b.hl) { const double pz = (s - b.lo) * vl; pmax = fmax(pmax, pz); pmin = fmin(pmin, pz); }
      if (b.hu) { const double pz = (b.hi - s) * vu; pmax = fmax(pmax, pz); pmin = fmin(pmin, pz); }
      const double lo_o = __ldg(A.lbg + lane * L::R + r), hi_o = __ldg(A.ubg + lane * L::R + r);
      if (lo_o > -1e19) viol = fmax(viol, lo_o - gu);
      if (hi_o < 1e19) viol = fmax(viol, gu - hi_o);
    };
    FOR_ROWS(st.X, body);
#pragma unroll
    for (int e = 0; e < 21; ++e) LQ(LQ_Q + e) = q66[e];
    LQ(LQ_NN + 0) = nn[0]; LQ(LQ_NN + 1) = nn[1]; LQ(LQ_NN + 2) = nn[2];
    LQ(LQ_ZERO) = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { LQ(LQ_QA + i) = qa[i]; LQ(LQ_QB + i) = ls ? 0.0 : qb[i]; LQ(LQ_QD + i) = ls ? 0.0 : qd[i]; }
    LQ(LQ_DD + 0) = st.cps * st.cth; LQ(LQ_DD + 1) = st.sps * st.cth; LQ(LQ_DD + 2) = st.sth;
    LQ(LQ_EE + 0) = dy.e03; LQ(LQ_EE + 1) = dy.e13; LQ(LQ_EE + 2) = dy.e23; LQ(LQ_EE + 3) = dy.e04; LQ(LQ_EE + 4) = dy.e14;
    LQ(LQ_Q + 21) = q22;
  }
  double lamn[8];
  adjoint(a, dy, act, lane, lamn);
  if (act) {
    double svt = 0.0, svp = 0.0, q33 = 0.0, q34 = 0.0, q44 = 0.0;
    if (hasu && !ls) {
      // curvature of T*v*d(theta,psi) weighted by the next-stage adjoint
      const double L0 = T * lamn[0], L1 = T * lamn[1], L2 = T * lamn[2], v = u[0];
      const double cc = st.cps * st.cth, sc = st.sps * st.cth, cs = st.cps * st.sth, ss = st.sps * st.sth;
      q33 = -v * (L0 * cc + L1 * sc + L2 * st.sth);
      q44 = -v * (L0 * cc + L1 * sc);
      q34 = v * (L0 * ss - L1 * cs);
      svt = -L0 * cs - L1 * ss + L2 * st.cth;
      svp = -L0 * sc + L1 * cc;
    }
    LQ(LQ_Q + 21) += q33; LQ(LQ_Q + 22) = q34; LQ(LQ_Q + 23) = q44;
    LQ(LQ_SV + 0) = svt; LQ(LQ_SV + 1) = svp;
    // controls: Sigma_x, gradient per unit mu, dual infeasibility, complementarity products
    double glx[6];
    glx[0] = T * (st.cps * st.cth * lamn[0] + st.sps * st.cth * lamn[1] + st.sth * lamn[2]);
#pragma unroll
    for (int r = 1; r < 6; ++r) glx[r] = T * lamn[r + 2];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double sig = 1.0, rb = 0.0;
      if (hasu) {
        const Bnd b = ctl_bounds(A, lane, i);
        const double zl = LV(LV_ZL + i), zu = LV(LV_ZU + i);
        const double sl = u[i] - b.lo, su = b.hi - u[i];
        if (ls) rb = zu - zl;
        else {
          const double il = b.hl ? rcp(sl) : 0.0, iu = b.hu ? rcp(su) : 0.0;
          sig = zl * il + zu * iu; rb = iu - il;
          if (b.hl && !b.hu) rb += A.o.kappa_d;
          if (b.hu && !b.hl) rb -= A.o.kappa_d;
        }
        du_l = fmax(du_l, fabs(glx[i] - zl + zu)); sumz += zl + zu;
        if (b.hl) { const double pz = sl * zl; pmax = fmax(pmax, pz); pmin = fmin(pmin, pz); }
        if (b.hu) { const double pz = su * zu; pmax = fmax(pmax, pz); pmin = fmin(pmin, pz); }
      }
      LQ(LQ_RD + i) = sig; LQ(LQ_RB + i) = rb;
    }
  }
  du_l = warp_max(du_l); pr_l = warp_max(pr_l); sumy = warp_sum(sumy); sumz = warp_sum(sumz);
  viol = warp_max(viol); pmax = warp_max(pmax); pmin = warp_min(pmin);
  if (lane == 0) {
    RES(R_F) = fsum; RES(R_DU) = du_l; RES(R_PR) = pr_l; RES(R_SUMY) = sumy; RES(R_SUMZ) = sumz;
    RES(R_VIOL) = viol; RES(R_PMAX) = pmax; RES(R_PMIN) = pmin;
  }
  __syncwarp();
}

// least-squares multipliers from the solved LS system: y = G dx + (v_U - v_L); zero if too large
template <class L>
__device__ __noinline__ void ph_lsy(const SolveArgs& A, int lane, bool ok) {
  const bool act = lane <= L::N;
  double ymax = 0.0;
  if (act && ok) {
    double X[8], dx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { X[i] = LV(LV_X + i); dx[i] = LV(LV_DX + i); }
    auto body = [&](const int r, const bool box, const int si, const double gu, double nx, double ny, double iD) {
      const double dc = RW(A_DC, r);
      const double gd = box ? dc * dx[si] : -dc * (nx * dx[0] + ny * dx[1]);
      const double yv = gd + (RW(A_VU, r) - RW(A_VL, r));
      RW(A_Y, r) = yv; ymax = fmax(ymax, fabs(yv));
    };
    FOR_ROWS(X, body);
  }
  ymax = warp_max(ymax);
  if (!ok || !(ymax <= A.o.constr_mult_init_max)) {
    if (act) {
#pragma unroll 1
      for (int r = 0; r < L::R; ++r) RW(A_Y, r) = 0.0;
    }
  }
  __syncwarp();
}

// step in the slacks, fraction-to-the-boundary limits, directional derivative of the barrier function.
// soc: second-order-correction direction (residual CSOC, controls DUS, output DS2).
template <class L>
__device__ __noinline__ void ph_dir(const SolveArgs& A, int lane, double mu, double tau, bool soc) {
  const bool act = lane <= L::N, hasu = lane < L::N;
  double tp = 0.0, dnum = 0.0, dden = 1.0, gbd = 0.0, theta = 0.0; bool nottiny = false;
  const double tt = A.o.tiny_step_tol, kd = A.o.kappa_d;
  auto dual_frac = [&](double z, double dz) {   // track max of -dz/z over dz < 0 as a fraction
    if (dz < 0.0 && -dz * dden > dnum * z) { dnum = -dz; dden = z; }
  };
  if (act) {
    double X[8], dx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { X[i] = LV(LV_X + i); dx[i] = LV(LV_DX + i); }
    auto body = [&](const int r, const bool box, const int si, const double gu, double nx, double ny, double iD) {
      const double dc = RW(A_DC, r), s = RW(A_S, r), il = RW(A_IL, r), iu = RW(A_IU, r), vl = RW(A_VL, r), vu = RW(A_VU, r);
      const bool hl = il > 0.0, hu = iu > 0.0;
      const double c = soc ? SOC(SOC_CSOC + r) : RW(A_G, r) - s;
      const double gd = box ? dc * dx[si] : -dc * (nx * dx[0] + ny * dx[1]);
      const double ds = gd + c;
      if (soc) SOC(SOC_DS2 + r) = ds; else RW(A_DS, r) = ds;
      tp = fmax(tp, fmax(-ds * il, ds * iu));
      if (hl) dual_frac(vl, (mu - vl * ds) * il - vl);
      if (hu) dual_frac(vu, (mu + vu * ds) * iu - vu);
      double beta = iu - il;
      if (hl && !hu) beta += kd;
      if (hu && !hl) beta -= kd;
      gbd += mu * beta * ds; theta += fabs(c);
      nottiny = nottiny || (fabs(ds) > tt * (1.0 + fabs(s)));
    };
    FOR_ROWS(X, body);
    if (hasu) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const Bnd b = ctl_bounds(A, lane, i);
        const double u = LV(LV_U + i), du = soc ? SOC(SOC_DUS + i) : LV(LV_DU + i), zl = LV(LV_ZL + i), zu = LV(LV_ZU + i);
        const double il = b.hl ? rcp(u - b.lo) : 0.0, iu = b.hu ? rcp(b.hi - u) : 0.0;
        tp = fmax(tp, fmax(-du * il, du * iu));
        if (b.hl) dual_frac(zl, (mu - zl * du) * il - zl);
        if (b.hu) dual_frac(zu, (mu + zu * du) * iu - zu);
        double beta = iu - il;
        if (b.hl && !b.hu) beta += kd;
        if (b.hu && !b.hl) beta -= kd;
        gbd += mu * beta * du;
        nottiny = nottiny || (fabs(du) > tt * (1.0 + fabs(u)));
      }
#pragma unroll
      for (int v = 0; v < 6; ++v) gbd += LV(LV_GL + v) * dx[cost_state(v)];
    }
  }
  tp = warp_max(tp);
  const double td = warp_max(dnum / dden);
  gbd = warp_sum(gbd); theta = warp_sum(theta);
  const bool any_nt = __any_sync(FULL, nottiny);
  if (lane == 0) {
    RES(R_APR) = tp > tau ? tau / tp : 1.0;             // alpha = min(1, tau / max ratio)
    RES(R_ADU) = td > tau ? tau / td : 1.0;
    if (!soc) { RES(R_GBD) = gbd; RES(R_THETA) = theta; RES(R_TINY) = any_nt ? 0.0 : 1.0; }
  }
  __syncwarp();
}

// trial point u + alpha*du, s + alpha*ds: objective, constraint violation, barrier pieces; residual into CT
template <class L>
__device__ __noinline__ void ph_trial(const SolveArgs& A, int lane, double alpha, bool soc, double df) {
  const Prob& pr = A.pr;
  const bool act = lane <= L::N, hasu = lane < L::N;
  double ut[6], lb = 0.0, dt = 0.0;
  double prod = 1.0; int cnt = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    ut[i] = 0.0;
    if (hasu) {
      ut[i] = fma(alpha, soc ? SOC(SOC_DUS + i) : LV(LV_DU + i), LV(LV_U + i));   // same expression as ph_accept
      const Bnd b = ctl_bounds(A, lane, i);
      if (b.hl) prod *= ut[i] - b.lo;
      if (b.hu) prod *= b.hi - ut[i];
      if (b.hl && !b.hu) dt += ut[i] - b.lo;
      if (b.hu && !b.hl) dt += b.hi - ut[i];
      if (i == 2 || i == 5) { lb += n_log(prod); prod = 1.0; }
    }
  }
  Stage st; rollout(pr, &PAR(0), ut, lane, st);
  const double l = hasu ? stage_cost(with_weights(pr, PAR(NPAR), PAR(NPAR + 1)), st.X, stage_target<L>(A, lane).x, stage_target<L>(A, lane).y) : 0.0;
  double th = 0.0;
  if (act) {
    auto body = [&](const int r, const bool box, const int si, const double gu, double nx, double ny, double iD) {
      const double dc = RW(A_DC, r);
      const double g = __dmul_rn(dc, gu);
      const double sv = fma(alpha, soc ? SOC(SOC_DS2 + r) : RW(A_DS, r), RW(A_S, r));
      const double ct = g - sv;
      SOC(SOC_CT + r) = ct; th += fabs(ct);
      const Bnd b = row_bounds<L>(A, lane, r, dc);
      if (b.hl) { prod *= sv - b.lo; ++cnt; }
      if (b.hu) { prod *= b.hi - sv; ++cnt; }
      if (b.hl && !b.hu) dt += sv - b.lo;
      if (b.hu && !b.hl) dt += b.hi - sv;
      if (cnt >= 4) { lb += n_log(prod); prod = 1.0; cnt = 0; }
    };
    FOR_ROWS(st.X, body);
    if (cnt) lb += n_log(prod);
  }
  const double fs = df * warp_sum(l);
  th = warp_sum(th); lb = warp_sum(lb); dt = warp_sum(dt);
  if (lane == 0) { RES(R_FT) = fs; RES(R_THT) = th; RES(R_LBT) = lb; RES(R_DTT) = dt; }
  __syncwarp();
}

// SOC right-hand side: c_soc <- a_soc * c_soc + c(trial);  q' = grad l + G^T((Sigma_s + dw) c_soc + mu beta)
template <class L>
__device__ __noinline__ void ph_socrhs(const SolveArgs& A, int lane, double a_soc, double mu, double dw, bool first) {
  if (lane <= L::N) {
    double X[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { X[i] = LV(LV_X + i); q[i] = 0.0; }
#pragma unroll
    for (int v = 0; v < 6; ++v) q[cost_state(v)] = LV(LV_GL + v);
    const double kd = A.o.kappa_d;
    auto body = [&](const int r, const bool box, const int si, const double gu, double nx, double ny, double iD) {
      const double dc = RW(A_DC, r), il = RW(A_IL, r), iu = RW(A_IU, r);
      const bool hl = il > 0.0, hu = iu > 0.0;
      const double cprev = first ? RW(A_G, r) - RW(A_S, r) : SOC(SOC_CSOC + r);
      const double cs = a_soc * cprev + SOC(SOC_CT + r);
      SOC(SOC_CSOC + r) = cs;
      double beta = iu - il;
      if (hl && !hu) beta += kd;
      if (hu && !hl) beta -= kd;
      const double yh = (RW(A_VL, r) * il + RW(A_VU, r) * iu + dw) * cs + mu * beta;
      if (box) q[si] += dc * yh; else { q[0] -= dc * yh * nx; q[1] -= dc * yh * ny; }
    };
    FOR_ROWS(X, body);
#pragma unroll
    for (int i = 0; i < 8; ++i) SOC(SOC_Q2 + i) = q[i];
  }
  __syncwarp();
}

// accept the trial point: primal step alpha, dual step a_du, kappa_Sigma reset, new reciprocal slacks
template <class L>
__device__ __noinline__ void ph_accept(const SolveArgs& A, int lane, double alpha, double a_du, double mu, double dw, bool soc) {
  const bool act = lane <= L::N, hasu = lane < L::N;
  const double ks = A.o.kappa_sigma, kd = A.o.kappa_d, iks = 1.0 / A.o.kappa_sigma;
  if (hasu) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const Bnd b = ctl_bounds(A, lane, i);
      const double u = LV(LV_U + i), du = soc ? SOC(SOC_DUS + i) : LV(LV_DU + i);
      double zl = LV(LV_ZL + i), zu = LV(LV_ZU + i);
      if (b.hl) zl += a_du * ((mu - zl * du) * rcp(u - b.lo) - zl);
      if (b.hu) zu += a_du * ((mu + zu * du) * rcp(b.hi - u) - zu);
      const double un = fma(alpha, du, u);
      if (b.hl) { const double i2 = rcp(un - b.lo); zl = fmax(fmin(zl, ks * mu * i2), mu * i2 * iks); }
      if (b.hu) { const double i2 = rcp(b.hi - un); zu = fmax(fmin(zu, ks * mu * i2), mu * i2 * iks); }
      LV(LV_U + i) = un; LV(LV_ZL + i) = zl; LV(LV_ZU + i) = zu;
    }
  }
  if (act) {
#pragma unroll 1
    for (int r = 0; r < L::R; ++r) {
      const double dc = RW(A_DC, r), s = RW(A_S, r), il = RW(A_IL, r), iu = RW(A_IU, r), y = RW(A_Y, r);
      const double ds = soc ? SOC(SOC_DS2 + r) : RW(A_DS, r);
      const bool hl = il > 0.0, hu = iu > 0.0;
      double vl = RW(A_VL, r), vu = RW(A_VU, r);
      double beta = iu - il;
      if (hl && !hu) beta += kd;
      if (hu && !hl) beta -= kd;
      const double dy = (vl * il + vu * iu + dw) * ds + (mu * beta - y);
      RW(A_Y, r) = y + alpha * dy;
      if (hl) vl += a_du * ((mu - vl * ds) * il - vl);
      if (hu) vu += a_du * ((mu + vu * ds) * iu - vu);
      const double sn = fma(alpha, ds, s);
      const Bnd b = row_bounds<L>(A, lane, r, dc);
      double il2 = 0.0, iu2 = 0.0;
      if (hl) { il2 = rcp(sn - b.lo); vl = fmax(fmin(vl, ks * mu * il2), mu * il2 * iks); }
      if (hu) { iu2 = rcp(b.hi - sn); vu = fmax(fmin(vu, ks * mu * iu2), mu * iu2 * iks); }
      RW(A_S, r) = sn; RW(A_VL, r) = vl; RW(A_VU, r) = vu; RW(A_IL, r) = il2; RW(A_IU, r) = iu2;
    }
  }
  __syncwarp();
}

// outputs: honour the original bounds, unscale multipliers, f and g at the returned point
template <class L>
__device__ __noinline__ void ph_output(const SolveArgs& A, int b, int lane, double df, int status, int iter) {
  const Prob& pr = A.pr; constexpr int N = L::N, R = L::R, S = L::S;
  const bool act = lane <= N, hasu = lane < N;
  constexpr int nw = NU * N, ng = R * S;
  double u[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) u[i] = 0.0;
  if (hasu) {
#pragma unroll
    for (int i = 0; i < 6; ++i) u[i] = fmin(fmax(LV(LV_U + i), __ldg(A.lbx + NU * lane + i)), __ldg(A.ubx + NU * lane + i));
    if (A.x) {
      double* xo = A.x + (size_t)b * nw + NU * lane;
#pragma unroll
      for (int i = 0; i < 6; ++i) xo[i] = u[i];
    }
    if (A.lam_x) {
      double* lo = A.lam_x + (size_t)b * nw + NU * lane;
      const double idf = 1.0 / df;
#pragma unroll
      for (int i = 0; i < 6; ++i) lo[i] = (LV(LV_ZU + i) - LV(LV_ZL + i)) * idf;
    }
  }
  Stage st; rollout(pr, &PAR(0), u, lane, st);
  const double l = hasu ? stage_cost(with_weights(pr, PAR(NPAR), PAR(NPAR + 1)), st.X, stage_target<L>(A, lane).x, stage_target<L>(A, lane).y) : 0.0;
  const double fu = warp_sum(l);
  if (lane == 0) {
    if (A.f) A.f[b] = fu;
    if (A.status) A.status[b] = status;
    if (A.iters) A.iters[b] = iter;
    if (A.iters_keep) A.iters_keep[b] = iter;
  }
  if (act) {
    if (A.g) {
      double* go = A.g + (size_t)b * ng + lane * R;
      auto body = [&](const int r, const bool box, const int si, const double gu, double nx, double ny, double iD) { go[r] = gu; };
      FOR_ROWS(st.X, body);
    }
    if (A.lam_g) {
      double* lo = A.lam_g + (size_t)b * ng + lane * R;
      const double idf = 1.0 / df;
#pragma unroll 1
      for (int r = 0; r < R; ++r) lo[r] = RW(A_Y, r) * RW(A_DC, r) * idf;
    }
  }
  if (A.step_p) {     // fused closed-loop shift of this instance (NMPC_TT.py:13-30): warm start, plant, target, FOV error
    double* uw = A.step_u + (size_t)b * nw;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double un = __shfl_down_sync(FULL, u[i], 1);          // control of the next stage
      if (lane < N - 1) uw[NU * lane + i] = un;                   // drop the first stage ...
      else if (lane == N - 1) uw[NU * lane + i] = u[i];           // ... and repeat the last (:20-23)
    }
    if (lane == 0)
      closed_loop_shift(pr.T, pr.hv, pr.hh, A.step_p + (size_t)b * NPAR, u, __ldg(A.step_vw + 2 * (size_t)b), __ldg(A.step_vw + 2 * (size_t)b + 1),
                        A.step_fov ? A.step_fov + 2 * (size_t)b : nullptr, A.step_err ? A.step_err + b : nullptr);
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
template <class L>
__device__ __noinline__ void solve_instance(const SolveArgs& A, double* ric, int b, int lane) {
  const Prob& pr = A.pr; const Opt& o = A.o;
  constexpr int S = L::S, R = L::R;
  constexpr int DX0 = L::LV0 + LV_DX * S, DU0 = L::LV0 + LV_DU * S, DUS0 = L::soc(3 * R), Q20 = L::soc(3 * R + 6);
  const double T = pr.T;
  unsigned long long n_fact = 0, n_ls = 0, n_soc = 0;

  ph_load<L>(A, b, lane);
  double df = 1.0;
  if (o.scaling) {
    df = ph_scaling<L>(A, lane);
    if (o.scaling == 3) df = 1.0;                                  // debug: constraint scaling only
    if (o.scaling == 2 && lane <= L::N) { for (int r = 0; r < R; ++r) RW(A_DC, r) = 1.0; }   // debug: objective only
    __syncwarp();
  }
  const int nzt = ph_start<L>(A, lane);
  ph_trial<L>(A, lane, 0.0, false, df);          // barrier pieces and constraint violation of the start
  double f = RES(R_FT), LB = RES(R_LBT), DT = RES(R_DTT);
  const double th0 = RES(R_THT);
  const double theta_max = o.theta_max_fact * fmax(1.0, th0), theta_min = o.theta_min_fact * fmax(1.0, th0);
  double mu = o.mu_init, tau = fmax(o.tau_min, 1.0 - mu);
  const double mu_floor = fmin(o.tol, o.compl_inf_tol) / (o.kappa_eps + 1.0);
  int nfilt = 0; double dw_last = 0.0;
  int iter = 0, status = NMPC_MAXITER_EXCEEDED, tiny_count = 0; bool tiny_flag = false, ls = true;
  constexpr int mtot = R * S;

  for (;;) {
    // Alignment point: the warps of a block start every IPM iteration together, so that they walk through the
    // same ~200 KB of phase code at the same time and share instruction-cache lines (unaligned warps thrash it:
    // `no_instruction` was 56 % of all stall cycles in v3).  Pure scheduling; results cannot depend on it.
    // (Aligning at more points per iteration -- before the factorisation, direction, line search, accept -- was
    // measured and is slower: each phase then lasts as long as its slowest warp; 43 -> 65 ms at B = 16384.)
    // (Measured and rejected: releasing the barrier with a quorum of 5-7 of 8 warps, and letting warps with long line
    // searches pay their arrival in advance and run out of phase -- both slower; strict lockstep wins.)
    align_warps(A.align_group, 1);
    ph_derivs<L>(A, lane, ls, df);
    if (ls) {   // least-squares multiplier start: (I + J^T J) t = -(grad_x L) - J^T (grad_s L),  y = J t + grad_s L
      ++n_fact;
      const bool ok = riccati_factor<L>(T, ric, A.ricmap, 1.0, 0.0, lane);
      if (ok) riccati_forward<L>(T, ric, false, lane, DX0, DU0);
      ph_lsy<L>(A, lane, ok);
      ls = false;
      continue;
    }
    // ---- optimality error (scaled) and termination
    const double du_inf = RES(R_DU), pr_inf = RES(R_PR), sumy = RES(R_SUMY), sumz = RES(R_SUMZ), viol = RES(R_VIOL);
    const double pmax = RES(R_PMAX), pmin = RES(R_PMIN);
    const double sd = fmax(o.s_max, (sumy + sumz) / (double)max(1, mtot + nzt)) / o.s_max;
    const double sc = fmax(o.s_max, sumz / (double)max(1, nzt)) / o.s_max;
    const double E0 = fmax(du_inf / sd, fmax(pr_inf, pmax / sc));
    if (!isfinite(E0) || !isfinite(f)) { status = NMPC_INVALID_NUMBER; break; }
    if (E0 <= o.tol && du_inf / df <= o.dual_inf_tol && viol <= o.constr_viol_tol && pmax / df <= o.compl_inf_tol) {
      status = NMPC_SOLVE_SUCCEEDED; break;
    }
    if (iter >= o.max_iter) { status = NMPC_MAXITER_EXCEEDED; break; }
    // ---- barrier parameter (monotone, with fast decrease)
    {
      auto emu = [&](double m) { return fmax(du_inf / sd, fmax(pr_inf, fmax(pmax - m, m - pmin) / sc)); };
      double Emu = emu(mu);
      while ((Emu <= o.kappa_eps * mu || tiny_flag) && mu > mu_floor) {
        mu = fmax(mu_floor, fmin(o.kappa_mu * mu, n_pow(mu, o.theta_mu)));
        tau = fmax(o.tau_min, 1.0 - mu); nfilt = 0; tiny_flag = false;
        Emu = emu(mu);
      }
      if (tiny_flag && mu <= mu_floor) { status = NMPC_STEP_TOO_SMALL; break; }
    }
    // ---- search direction with inertia correction
    const unsigned long long ls_before = n_ls;
    double dw = 0.0; bool ok = false;
    for (;;) {
      ++n_fact;
      ok = riccati_factor<L>(T, ric, A.ricmap, mu, dw, lane);
      if (ok) break;
      if (dw == 0.0) dw = (dw_last == 0.0) ? o.dw_init : fmax(o.dw_min, dw_last * o.dw_dec);
      else dw = (dw_last == 0.0 || 1e5 * dw_last < dw) ? o.dw_inc_first * dw : o.dw_inc * dw;
      if (dw > o.dw_max) break;
    }
    if (!ok) { status = NMPC_PERTURBATION_FAILED; break; }
    if (dw > 0.0) dw_last = dw;
    riccati_forward<L>(T, ric, false, lane, DX0, DU0);
    ph_dir<L>(A, lane, mu, tau, false);
    const double a_pr_max = RES(R_APR); double a_du = RES(R_ADU);
    const double gbd = RES(R_GBD), theta = RES(R_THETA);
    const bool tiny = RES(R_TINY) != 0.0 && theta <= 1e-4;
    const double phi = f - mu * LB + o.kappa_d * mu * DT;
    // ---- filter line search
    const double pw_gbd = gbd < 0.0 ? n_pow(-gbd, o.s_phi) : 0.0, pw_th = n_pow(theta, o.s_theta);
    auto is_ftype = [&](double a) { return gbd < 0.0 && a * pw_gbd > o.delta * pw_th; };
    auto armijo = [&](double a, double ph_t) { return cmp_le(ph_t - phi, o.eta_phi * a * gbd, phi); };
    auto acceptable = [&](double a_test, double th_, double ph_) {
      if (!isfinite(th_) || !isfinite(ph_)) return false;
      if (th_ > theta_max) return false;
      bool acc;
      if (a_test > 0.0 && is_ftype(a_test) && theta <= theta_min) acc = armijo(a_test, ph_);
      else acc = cmp_le(th_, (1.0 - o.gamma_theta) * theta, theta) || cmp_le(ph_ - phi, -o.gamma_phi * theta, phi);
      if (!acc) return false;
      for (int e = 0; e < nfilt; ++e) if (!(th_ < smem[L::FILT0 + 2 * e] || ph_ < smem[L::FILT0 + 2 * e + 1])) return false;
      return true;
    };
    double amin = o.gamma_theta;
    if (gbd < 0.0) {
      amin = fmin(o.gamma_theta, o.gamma_phi * theta / (-gbd));
      if (theta <= theta_min) amin = fmin(amin, o.delta * pw_th / pw_gbd);
    }
    amin *= o.alpha_min_frac;
    double alpha = a_pr_max, alpha_test = a_pr_max, phi_t = 0.0, a_soc = 0.0, th_prev = 0.0;
    bool accepted = false, used_soc = false, first = true;
    int soc_left = 0;      // > 0: the next trial is a second-order-correction trial
    if (tiny) { ++tiny_count; tiny_flag = true; } else tiny_count = 0;
    for (;;) {
      const bool soc_trial = soc_left > 0;
      const double a_try = soc_trial ? a_soc : alpha;
      ph_trial<L>(A, lane, a_try, soc_trial, df);
      ++n_ls;
      const double th_t = RES(R_THT);
      phi_t = RES(R_FT) - mu * RES(R_LBT) + o.kappa_d * mu * RES(R_DTT);
      if (tiny) { accepted = true; break; }
      if (!soc_trial) alpha_test = alpha;
      if (acceptable(alpha, th_t, phi_t)) {
        accepted = true;
        if (soc_trial) { used_soc = true; alpha = a_soc; ++n_soc; }
        break;
      }
      bool next_soc = false;
      if (soc_trial) {
        --soc_left;
        next_soc = soc_left > 0 && th_t <= o.kappa_soc * th_prev;
        if (!next_soc) soc_left = 0;
      } else if (first && o.max_soc > 0 && th_t >= theta && isfinite(th_t)) {
        soc_left = o.max_soc; next_soc = true;
      }
      if (next_soc) {
        const bool first_soc = !soc_trial;
        th_prev = th_t;
        ph_socrhs<L>(A, lane, first_soc ? alpha : a_soc, mu, dw, first_soc);
        riccati_resolve<L>(T, Q20, ric, mu, lane);
        riccati_forward<L>(T, ric, true, lane, DX0, DUS0);
        ph_dir<L>(A, lane, mu, tau, true);
        a_soc = RES(R_APR);
        continue;
      }
      first = false;
      alpha *= o.alpha_red;
      if (!(alpha > amin)) break;
    }
    if (tiny && tiny_count >= 2 && mu <= mu_floor) { status = NMPC_STEP_TOO_SMALL; break; }
    if (!accepted) { status = NMPC_RESTORATION_NEEDED; break; }
    // ---- filter augmentation
    if (!tiny && !(is_ftype(alpha_test) && armijo(alpha_test, phi_t))) {
      if (lane == 0) {
        if (nfilt == FILT_CAP) for (int e = 0; e < 2 * (FILT_CAP - 1); ++e) smem[L::FILT0 + e] = smem[L::FILT0 + e + 2];   // drop the oldest
        const int at = nfilt == FILT_CAP ? FILT_CAP - 1 : nfilt;
        smem[L::FILT0 + 2 * at] = (1.0 - o.gamma_theta) * theta; smem[L::FILT0 + 2 * at + 1] = phi - o.gamma_phi * theta;
      }
      if (nfilt < FILT_CAP) ++nfilt;
      __syncwarp();
    }
    if (used_soc) a_du = RES(R_ADU);
    if (A.dbg && lane == 0 && iter < A.dbg_rows) {
      double* Lg = A.dbg + ((size_t)b * A.dbg_rows + iter) * 8;
      Lg[0] = mu; Lg[1] = f / df; Lg[2] = pr_inf; Lg[3] = du_inf; Lg[4] = dw; Lg[5] = alpha; Lg[6] = a_du; Lg[7] = (double)(n_ls - ls_before);
    }
    ph_accept<L>(A, lane, alpha, a_du, mu, dw, used_soc);
    f = RES(R_FT); LB = RES(R_LBT); DT = RES(R_DTT);
    ++iter;
  }
  ph_output<L>(A, b, lane, df, status, iter);
  if (lane == 0 && A.stats) { atomicAdd(&A.stats[0], n_fact); atomicAdd(&A.stats[1], n_ls); atomicAdd(&A.stats[2], n_soc); }
}

// Longest-first fetch order for the NEXT call on this handle from this call's iteration counts: counting sort,
// descending, by the warp that finished the call's last instance (its shared-memory slice is free; the order inside a
// bin is arbitrary -- results do not depend on it).
template <class L>
__device__ __noinline__ void next_order(const SolveArgs& A, int lane) {
  int* bin = reinterpret_cast<int*>(&smem[0]);          // 256 ints
  for (int i = lane; i < 256; i += 32) bin[i] = 0;
  __syncwarp();
  for (int i = lane; i < A.B; i += 32) atomicAdd(&bin[min(max(__ldcg(A.iters_keep + i), 0), 255)], 1);
  __syncwarp();
  if (lane == 0) {
    int acc = 0;
    for (int v = 255; v >= 0; --v) { const int c = bin[v]; bin[v] = acc; acc += c; }
  }
  __syncwarp();
  for (int i = lane; i < A.B; i += 32) A.order_out[atomicAdd(&bin[min(max(__ldcg(A.iters_keep + i), 0), 255)], 1)] = i;
  __syncwarp();
}

// persistent kernel: Lay::WPB warps per block (one block per SM), one instance per warp at a time, instances from
// an atomic work queue (optionally in a caller-given order)
template <int N_, int NOBS_>
__global__ void __launch_bounds__(32 * Lay<N_, NOBS_>::WPB, 1) nmpc_ipm_kernel(const SolveArgs A) {
  using L = Lay<N_, NOBS_>;
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) {      // the next call on this handle starts from clean counters
    *A.counter_next = 0; *A.done_next = 0; A.stats_next[0] = 0; A.stats_next[1] = 0; A.stats_next[2] = 0;
  }
  double* ric = A.ric + (size_t)(blockIdx.x * L::WPB + (threadIdx.x >> 5)) * A.ric_stride;
  for (;;) {
    int q = 0;
    if (lane == 0) q = atomicAdd(A.counter, 1);
    q = __shfl_sync(FULL, q, 0);
    if (q >= A.B) break;
    const int b = A.order ? A.order[q] : q;
    solve_instance<L>(A, ric, b, lane);
    __syncwarp();
    int fin = 0;
    if (lane == 0) { __threadfence(); fin = atomicAdd(A.done, 1); }      // this instance's outputs are visible before the count
    fin = __shfl_sync(FULL, fin, 0);
    if (fin == A.B - 1 && A.order_out) next_order<L>(A, lane);           // last instance of the call
  }
  while (align_warps(A.align_group, 0)) {}   // out of work: keep matching the alignment barrier until the group is done
}

#undef LV
#undef RW
#undef LQ
#undef SOC
#undef RES
#undef PAR
#undef smem

}  // namespace nmpc
