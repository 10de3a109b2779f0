// TEST INFRASTRUCTURE ONLY (oracle).  Nothing under mpc-implementation_b200/ may include,
// link or call this file; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// --impl reference legs use it, as the checker or as the timed CPU baseline.
//
// PARITY UNPINNED.  The reference (devsonni/MPC-Implementation) delegates all arithmetic of
// its hot path -- `sol = solver(x0,lbx,ubx,lbg,ubg,p)`, Python/NMPC_TT.py:358-365 -- to
// CasADi (un-pinned; the MATLAB twins say "CasADi v3.5.5", MATLAB/Race track 2/NMPC_TT.m:2),
// which bundles IPOPT + MUMPS.  None of those is vendored in /root/reference or installable
// here, and the reference holds no tests / golden vectors / recorded outputs.  This file
// therefore restates
//   (1) the NLP exactly as the scripts build it:
//         dynamics + Euler rollout ..... Python/NMPC_TT.py:139-148, :160-167
//         objective ..................... Python/NMPC_TT.py:193-221 (literal a,b,A,B,C,X_E,Y_E form)
//         constraint rows ............... Python/NMPC_TT.py:234-244; Race Track 2.py:247-264
//       with first and second derivatives produced by forward-mode AD "jets" applied to the
//       literal formulas (the role CasADi's AD plays), assembled into the dense single-shooting
//       gradient / Jacobian / Lagrangian Hessian IPOPT would be handed;
//   (2) IPOPT's published algorithm (Waechter & Biegler, Math. Prog. 106 (2006), and the
//       documented option defaults) with the options the scripts set, NMPC_TT.py:257-265:
//       slack reformulation, bound relaxation 1e-8, gradient-based scaling (max grad 100),
//       bound_push/frac 1e-2 start, least-squares multiplier start, monotone barrier update
//       (mu0=0.1, kappa_mu=0.2, theta_mu=1.5, kappa_eps=10), fraction-to-boundary
//       tau=max(0.99,1-mu), inertia correction (1e-4, x100/x8, /3), filter line search with
//       second-order corrections (max_soc=4), kappa_Sigma=1e10 multiplier reset, scaled
//       optimality-error termination tol=1e-8, max_iter=100.
//       Linear algebra: the augmented system is reduced by eliminating slacks and multipliers
//       and solved by a dense Cholesky of the n_w x n_w condensed matrix (inertia of the
//       augmented matrix is correct  <=>  that matrix is positive definite).
//       Round 2 added what IPOPT does when that line search does not simply succeed, restated
//       from memory of IPOPT 3.12's sources [3P]: filter details (Compare_le slack, dominated
//       entries, obj_max_inc, filter resets), tiny-step handling, the watchdog, the soft
//       restoration phase, the restoration phase proper (MinC_1NrmRestorationPhase with n, p
//       eliminated; RestoFilterConvergenceCheck; multipliers after restoration) and the slack
//       safeguard -- with IPOPT's own exits (Maximum_Iterations_Exceeded, Restoration_Failed,
//       Infeasible_Problem_Detected, Search_Direction_Becomes_Too_Small).
//   Not restated: iterative refinement / residual-ratio heuristics of PDFullSpaceSolver, the
//   degeneracy heuristics of PDPerturbationHandler, expect_infeasible_problem, adaptive mu.
//   Independent check of the main loop: oracle/ipm_fullspace.py (full-space KKT system with a
//   symmetric-indefinite factorisation and inertia count, derivatives by torch.autograd of the
//   literal formulas) must reproduce this file's iterates (tests/test_oracle_solve.py).
//
// Build:  g++ -O3 -march=native -shared -fPIC -o oracle/_build/libnmpc_oracle.so oracle/nmpc_oracle.cpp -lpthread
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int NX = 8, NU = 6, NPAR = 11;
constexpr double INF = std::numeric_limits<double>::infinity();
constexpr double EPS = std::numeric_limits<double>::epsilon();

// ------------------------------------------------------------------------------------------
// forward-mode second-order AD
// ------------------------------------------------------------------------------------------
template <int NV>
struct Jet {
  double v;
  double g[NV];
  double h[NV][NV];
  Jet() : v(0) { std::memset(g, 0, sizeof g); std::memset(h, 0, sizeof h); }
  explicit Jet(double c) : v(c) { std::memset(g, 0, sizeof g); std::memset(h, 0, sizeof h); }
  static Jet var(double val, int idx) { Jet j(val); j.g[idx] = 1.0; return j; }
};

template <int NV>
Jet<NV> chain(const Jet<NV>& a, double f, double d1, double d2) {
  Jet<NV> r;
  r.v = f;
  for (int i = 0; i < NV; ++i) r.g[i] = d1 * a.g[i];
  for (int i = 0; i < NV; ++i)
    for (int j = 0; j < NV; ++j) r.h[i][j] = d1 * a.h[i][j] + d2 * a.g[i] * a.g[j];
  return r;
}
template <int NV> Jet<NV> operator+(const Jet<NV>& a, const Jet<NV>& b) {
  Jet<NV> r; r.v = a.v + b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] + b.g[i];
  for (int i = 0; i < NV; ++i) for (int j = 0; j < NV; ++j) r.h[i][j] = a.h[i][j] + b.h[i][j];
  return r;
}
template <int NV> Jet<NV> operator-(const Jet<NV>& a, const Jet<NV>& b) {
  Jet<NV> r; r.v = a.v - b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] - b.g[i];
  for (int i = 0; i < NV; ++i) for (int j = 0; j < NV; ++j) r.h[i][j] = a.h[i][j] - b.h[i][j];
  return r;
}
template <int NV> Jet<NV> operator*(const Jet<NV>& a, const Jet<NV>& b) {
  Jet<NV> r; r.v = a.v * b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] * b.v + a.v * b.g[i];
  for (int i = 0; i < NV; ++i)
    for (int j = 0; j < NV; ++j)
      r.h[i][j] = a.h[i][j] * b.v + a.v * b.h[i][j] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  return r;
}
template <int NV> Jet<NV> operator+(const Jet<NV>& a, double c) { Jet<NV> r = a; r.v += c; return r; }
template <int NV> Jet<NV> operator-(const Jet<NV>& a, double c) { Jet<NV> r = a; r.v -= c; return r; }
template <int NV> Jet<NV> operator-(double c, const Jet<NV>& a) { return Jet<NV>(c) - a; }
template <int NV> Jet<NV> operator*(double c, const Jet<NV>& a) { return chain(a, c * a.v, c, 0.0); }
template <int NV> Jet<NV> operator*(const Jet<NV>& a, double c) { return c * a; }
template <int NV> Jet<NV> operator/(const Jet<NV>& a, double c) { return (1.0 / c) * a; }
template <int NV> Jet<NV> recip(const Jet<NV>& a) { double i = 1.0 / a.v; return chain(a, i, -i * i, 2 * i * i * i); }
template <int NV> Jet<NV> operator/(const Jet<NV>& a, const Jet<NV>& b) { return a * recip(b); }
template <int NV> Jet<NV> operator/(double c, const Jet<NV>& b) { return c * recip(b); }
template <int NV> Jet<NV> sq(const Jet<NV>& a) { return chain(a, a.v * a.v, 2 * a.v, 2.0); }
template <int NV> Jet<NV> jsqrt(const Jet<NV>& a) { double s = std::sqrt(a.v); return chain(a, s, 0.5 / s, -0.25 / (s * a.v)); }
template <int NV> Jet<NV> jsin(const Jet<NV>& a) { double s = std::sin(a.v), c = std::cos(a.v); return chain(a, s, c, -s); }
template <int NV> Jet<NV> jcos(const Jet<NV>& a) { double s = std::sin(a.v), c = std::cos(a.v); return chain(a, c, -s, -c); }
template <int NV> Jet<NV> jtan(const Jet<NV>& a) { double t = std::tan(a.v); double d = 1 + t * t; return chain(a, t, d, 2 * t * d); }

// ------------------------------------------------------------------------------------------
// problem description
// ------------------------------------------------------------------------------------------
// model 0: the scripts' UAV with a gimballed camera, 8 states / 6 controls, p = [state(8); x_t; y_t; theta_t]   NMPC_TT.py:105-154
// model 1: the gimbal-less tracker of MATLAB/Dynamic Obstacles/NMPC_TT.m:25-35 -- states (x,y,z,theta,psi), controls
//          (v, omega_2, omega_3), p = [state(5); x_t; y_t; theta_t], cost = horizontal distance only (:100-104), rows
//          [z, theta] per stage (:107-111).  Its dynamics are the first five rows of model 0's, so the rollout and the
//          sensitivities below run on the 8-state arrays with the camera states held at zero.
struct Spec {
  double T; int N; int n_obs; double w1, w2, vfov, hfov; int model = 0;
  int nu() const { return model ? 3 : NU; }
  int nlin() const { return model ? 2 : 5; }
  int npar() const { return model ? 8 : NPAR; }
  int rows() const { return nlin() + n_obs; }
  int nw() const { return nu() * N; }
  int ng() const { return rows() * (N + 1); }
};

// stage cost, literal NMPC_TT.py:209-220.  variables: (x, y, z, X5, X6, X7) -> state idx {0,1,2,5,6,7}
const int COST_IDX[6] = {0, 1, 2, 5, 6, 7};
template <class S>
S stage_cost_literal(const Spec& sp, const S& x, const S& y, const S& z, const S& X5, const S& X6, const S& X7,
                     double xt, double yt,
                     S (*Tan)(const S&), S (*Sin)(const S&), S (*Cos)(const S&), S (*Sqrt)(const S&), S (*Sq)(const S&)) {
  const double VF = sp.vfov, HF = sp.hfov;
  S a = (z * Tan(X6 + VF / 2) - z * Tan(X6 - VF / 2)) / 2.0;
  S b = (z * Tan(X5 + HF / 2) - z * Tan(X5 - HF / 2)) / 2.0;
  S c7 = Cos(X7), s7 = Sin(X7);
  S ia2 = 1.0 / Sq(a), ib2 = 1.0 / Sq(b);
  S A = Sq(c7) * ia2 + Sq(s7) * ib2;
  S B = 2.0 * c7 * s7 * (ia2 - ib2);
  S C = Sq(s7) * ia2 + Sq(c7) * ib2;
  S XE = x + a + z * Tan(X6 - VF / 2);
  S YE = y + b + z * Tan(X5 - HF / 2);
  S ex = xt - XE, ey = yt - YE;
  S dist = Sqrt(Sq(x - xt) + Sq(y - yt));
  return sp.w1 * dist + sp.w2 * ((A * Sq(ex) + B * ey * ex + C * Sq(ey)) - 1.0);
}
// plain-double instantiation helpers
struct D {
  double v; D() : v(0) {} D(double x) : v(x) {}
};
inline D operator+(D a, D b) { return a.v + b.v; } inline D operator-(D a, D b) { return a.v - b.v; }
inline D operator*(D a, D b) { return a.v * b.v; } inline D operator/(D a, D b) { return a.v / b.v; }
inline D operator+(D a, double b) { return a.v + b; } inline D operator-(D a, double b) { return a.v - b; }
inline D operator-(double a, D b) { return a - b.v; } inline D operator*(double a, D b) { return a * b.v; }
inline D operator*(D a, double b) { return a.v * b; } inline D operator/(D a, double b) { return a.v / b; }
inline D operator/(double a, D b) { return a / b.v; }
D dtan(const D& a) { return std::tan(a.v); } D dsin(const D& a) { return std::sin(a.v); }
D dcos(const D& a) { return std::cos(a.v); } D dsqrt(const D& a) { return std::sqrt(a.v); }
D dsq(const D& a) { return a.v * a.v; }
using J6 = Jet<6>;
J6 j6tan(const J6& a) { return jtan(a); } J6 j6sin(const J6& a) { return jsin(a); }
J6 j6cos(const J6& a) { return jcos(a); } J6 j6sqrt(const J6& a) { return jsqrt(a); } J6 j6sq(const J6& a) { return sq(a); }

// ------------------------------------------------------------------------------------------
// one NLP instance: function and dense derivative evaluation
// ------------------------------------------------------------------------------------------
struct Instance {
  Spec sp;
  const double* p;      // 11 (model 1: 8)
  const double* obs;    // n_obs x 3 : cx, cy, r_uav + r_obs
  int nw, ng, rows, nu, nlin, pt;
  // work (valid after eval_point)
  std::vector<double> X;        // (N+1) x 8
  std::vector<double> Sx;       // (N+1) x 8 x nw  sensitivities dX_k/dw
  std::vector<J6> cost;         // N
  std::vector<Jet<2>> obsj;     // (N+1) x n_obs
  std::vector<Jet<3>> dyn;      // N x 3   (theta, psi, v) -> T*rhs rows 0..2

  // optional per-stage predicted target [N][2] (the reference keeps (x_t, y_t) = p[8:10] over the horizon)
  const double* tgt = nullptr;
  double xt(int k) const { return tgt ? tgt[2 * k] : p[pt]; }
  double yt(int k) const { return tgt ? tgt[2 * k + 1] : p[pt + 1]; }

  Instance(const Spec& s, const double* p_, const double* obs_) : sp(s), p(p_), obs(obs_) {
    nw = sp.nw(); ng = sp.ng(); rows = sp.rows(); nu = sp.nu(); nlin = sp.nlin(); pt = sp.model ? 5 : 8;
    X.resize((sp.N + 1) * NX);
  }
  // stage cost: model 0 = NMPC_TT.py:209-220, model 1 = MATLAB/Dynamic Obstacles/NMPC_TT.m:100-104 (distance only, no weights)
  template <class S>
  S cost_at(const S& x, const S& y, const S& z, const S& X5, const S& X6, const S& X7, int k,
            S (*Tan)(const S&), S (*Sin)(const S&), S (*Cos)(const S&), S (*Sqrt)(const S&), S (*Sq)(const S&)) const {
    if (sp.model) return Sqrt(Sq(x - xt(k)) + Sq(y - yt(k)));
    return stage_cost_literal<S>(sp, x, y, z, X5, X6, X7, xt(k), yt(k), Tan, Sin, Cos, Sqrt, Sq);
  }

  void rollout(const double* w) {
    for (int i = 0; i < NX; ++i) X[i] = i < (sp.model ? 5 : NX) ? p[i] : 0.0;
    for (int k = 0; k < sp.N; ++k) {
      const double* st = &X[k * NX]; double u[NU] = {0, 0, 0, 0, 0, 0}; double* nx = &X[(k + 1) * NX];
      for (int i = 0; i < nu; ++i) u[i] = w[nu * k + i];
      double th = st[3], ps = st[4], v = u[0];
      nx[0] = st[0] + sp.T * (v * std::cos(ps) * std::cos(th));
      nx[1] = st[1] + sp.T * (v * std::sin(ps) * std::cos(th));
      nx[2] = st[2] + sp.T * (v * std::sin(th));
      nx[3] = st[3] + sp.T * u[1];
      nx[4] = st[4] + sp.T * u[2];
      nx[5] = st[5] + sp.T * u[3];
      nx[6] = st[6] + sp.T * u[4];
      nx[7] = st[7] + sp.T * u[5];
    }
  }
  // f and g only (line-search trial points)
  double eval_fg(const double* w, double* g) {
    rollout(w);
    double f = 0;
    for (int k = 0; k < sp.N; ++k) {
      const double* st = &X[k * NX];
      f += cost_at<D>(st[0], st[1], st[2], st[5], st[6], st[7], k, dtan, dsin, dcos, dsqrt, dsq).v;
    }
    for (int k = 0; k <= sp.N; ++k) {
      const double* st = &X[k * NX]; double* gk = g + k * rows;
      static const int LIN[5] = {2, 3, 5, 6, 7};
      for (int i = 0; i < nlin; ++i) gk[i] = st[LIN[i]];
      for (int j = 0; j < sp.n_obs; ++j) {
        double dx = st[0] - obs[3 * j], dy = st[1] - obs[3 * j + 1];
        gk[nlin + j] = -std::sqrt(dx * dx + dy * dy) + obs[3 * j + 2];
      }
    }
    return f;
  }
  // jets + sensitivities at w (rollout must correspond to w)
  void eval_derivs(const double* w) {
    rollout(w);
    const int N = sp.N;
    cost.assign(N, J6()); obsj.assign((N + 1) * sp.n_obs, Jet<2>()); dyn.assign(N * 3, Jet<3>());
    for (int k = 0; k < N; ++k) {
      const double* st = &X[k * NX];
      J6 v[6];
      for (int i = 0; i < 6; ++i) v[i] = J6::var(st[COST_IDX[i]], i);
      cost[k] = cost_at<J6>(v[0], v[1], v[2], v[3], v[4], v[5], k, j6tan, j6sin, j6cos, j6sqrt, j6sq);
      Jet<3> th = Jet<3>::var(st[3], 0), ps = Jet<3>::var(st[4], 1), vv = Jet<3>::var(w[nu * k], 2);
      dyn[k * 3 + 0] = sp.T * (vv * jcos(ps) * jcos(th));
      dyn[k * 3 + 1] = sp.T * (vv * jsin(ps) * jcos(th));
      dyn[k * 3 + 2] = sp.T * (vv * jsin(th));
    }
    for (int k = 0; k <= N; ++k) {
      const double* st = &X[k * NX];
      for (int j = 0; j < sp.n_obs; ++j) {
        Jet<2> x = Jet<2>::var(st[0], 0), y = Jet<2>::var(st[1], 1);
        obsj[k * sp.n_obs + j] = obs[3 * j + 2] - jsqrt(sq(x - obs[3 * j]) + sq(y - obs[3 * j + 1]));
      }
    }
    // sensitivities  Sx_{k+1} = A_k Sx_k + B_k E_k
    Sx.assign((size_t)(N + 1) * NX * nw, 0.0);
    for (int k = 0; k < N; ++k) {
      const double* S0 = &Sx[(size_t)k * NX * nw]; double* S1 = &Sx[(size_t)(k + 1) * NX * nw];
      for (int c = 0; c < nu * k; ++c) {   // only columns of earlier stages are non-zero
        for (int r = 0; r < NX; ++r) S1[r * nw + c] = S0[r * nw + c];
        for (int r = 0; r < 3; ++r)
          S1[r * nw + c] += dyn[k * 3 + r].g[0] * S0[3 * nw + c] + dyn[k * 3 + r].g[1] * S0[4 * nw + c];
      }
      int c0 = nu * k;
      for (int r = 0; r < 3; ++r) S1[r * nw + c0] = dyn[k * 3 + r].g[2];
      for (int i = 1; i < nu; ++i) S1[(2 + i) * nw + c0 + i] = sp.T;   // rows 3..7 <- controls 1..5
    }
  }
  void grad_f(double* grad) const {
    std::fill(grad, grad + nw, 0.0);
    for (int k = 1; k < sp.N; ++k) {   // stage 0 is constant in w
      const double* S = &Sx[(size_t)k * NX * nw];
      for (int i = 0; i < 6; ++i) {
        double gi = cost[k].g[i]; const double* row = S + COST_IDX[i] * nw;
        for (int c = 0; c < nu * k; ++c) grad[c] += gi * row[c];
      }
    }
  }
  // J : ng x nw, row-major
  void jac_g(double* J) const {
    std::fill(J, J + (size_t)ng * nw, 0.0);
    static const int LIN[5] = {2, 3, 5, 6, 7};
    for (int k = 1; k <= sp.N; ++k) {
      const double* S = &Sx[(size_t)k * NX * nw];
      for (int i = 0; i < nlin; ++i) {
        double* Jr = J + (size_t)(k * rows + i) * nw; const double* row = S + LIN[i] * nw;
        for (int c = 0; c < nu * k; ++c) Jr[c] = row[c];
      }
      for (int j = 0; j < sp.n_obs; ++j) {
        double* Jr = J + (size_t)(k * rows + nlin + j) * nw; const Jet<2>& o = obsj[k * sp.n_obs + j];
        for (int c = 0; c < nu * k; ++c) Jr[c] = o.g[0] * S[c] + o.g[1] * S[nw + c];
      }
    }
  }
  // W = Hess_w ( sigma f + lam^T g ), nw x nw dense symmetric
  void hess_l(double sigma, const double* lam, double* W) const {
    const int N = sp.N;
    std::fill(W, W + (size_t)nw * nw, 0.0);
    // adjoint of the Lagrangian through the rollout: lamx_k = dL/dX_k (total)
    std::vector<double> lamx((N + 1) * NX, 0.0);
    static const int LIN[5] = {2, 3, 5, 6, 7};
    for (int k = N; k >= 1; --k) {
      double* l = &lamx[k * NX];
      if (k < N) for (int i = 0; i < 6; ++i) l[COST_IDX[i]] += sigma * cost[k].g[i];
      for (int i = 0; i < nlin; ++i) l[LIN[i]] += lam[k * rows + i];
      for (int j = 0; j < sp.n_obs; ++j) {
        l[0] += lam[k * rows + nlin + j] * obsj[k * sp.n_obs + j].g[0];
        l[1] += lam[k * rows + nlin + j] * obsj[k * sp.n_obs + j].g[1];
      }
      if (k < N) {   // + A_k^T lamx_{k+1}
        const double* ln = &lamx[(k + 1) * NX];
        for (int i = 0; i < NX; ++i) l[i] += ln[i];
        for (int r = 0; r < 3; ++r) { l[3] += dyn[k * 3 + r].g[0] * ln[r]; l[4] += dyn[k * 3 + r].g[1] * ln[r]; }
      }
    }
    // stage Hessians in z=(X_k (8), U_k (6)) and projection W += Z^T H Z
    std::vector<double> Z(14 * (size_t)nw), HZ(14 * (size_t)nw);
    for (int k = 0; k <= N; ++k) {
      double H[14][14]; std::memset(H, 0, sizeof H);
      if (k >= 1 && k < N)
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) H[COST_IDX[i]][COST_IDX[j]] += sigma * cost[k].h[i][j];
      if (k >= 1)
        for (int j = 0; j < sp.n_obs; ++j) {
          const Jet<2>& o = obsj[k * sp.n_obs + j]; double l = lam[k * rows + nlin + j];
          for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) H[a][b] += l * o.h[a][b];
        }
      if (k < N) {   // dynamics curvature weighted by the next-stage adjoint; vars (theta=3, psi=4, v=8)
        static const int DI[3] = {3, 4, 8};
        const double* ln = &lamx[(k + 1) * NX];
        for (int r = 0; r < 3; ++r)
          for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H[DI[a]][DI[b]] += ln[r] * dyn[k * 3 + r].h[a][b];
      }
      int ncol = std::min(nw, nu * (k + 1));   // Z's non-zero columns: controls of stages <= k
      std::fill(Z.begin(), Z.end(), 0.0);
      const double* S = &Sx[(size_t)k * NX * nw];
      for (int r = 0; r < NX; ++r) for (int c = 0; c < nu * k; ++c) Z[(size_t)r * nw + c] = S[(size_t)r * nw + c];
      if (k < N) for (int i = 0; i < nu; ++i) Z[(size_t)(NX + i) * nw + nu * k + i] = 1.0;
      for (int r = 0; r < 14; ++r) {
        double* hz = &HZ[(size_t)r * nw]; std::fill(hz, hz + ncol, 0.0);
        for (int q = 0; q < 14; ++q) { double h = H[r][q]; if (h == 0.0) continue; const double* zq = &Z[(size_t)q * nw];
          for (int c = 0; c < ncol; ++c) hz[c] += h * zq[c]; }
      }
      for (int q = 0; q < 14; ++q) {
        const double* zq = &Z[(size_t)q * nw]; const double* hz = &HZ[(size_t)q * nw];
        for (int a = 0; a < ncol; ++a) { double za = zq[a]; if (za == 0.0) continue; double* Wr = W + (size_t)a * nw;
          for (int c = 0; c < ncol; ++c) Wr[c] += za * hz[c]; }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// dense Cholesky
// ------------------------------------------------------------------------------------------
bool cholesky(double* A, int n) {   // lower, in place; false if not positive definite
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d); A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j]; const double* ai = A + (size_t)i * n; const double* aj = A + (size_t)j * n;
      for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  return true;
}
void chol_solve(const double* L, int n, double* b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * b[k]; b[i] = s / L[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * b[k]; b[i] = s / L[(size_t)i * n + i]; }
}


// ------------------------------------------------------------------------------------------
// IPOPT-algorithm interior point method
//
// [3P] Everything below restates IPOPT (3.12 line; the version CasADi 3.5.5 bundles) from its paper and
// documented defaults -- the class names in the comments say which IPOPT component a block stands for:
//   Algo::optimize ................. IpoptAlgorithm::Optimize
//   Algo::check_convergence ........ OptimalityErrorConvergenceCheck / RestoFilterConvergenceCheck
//   Algo::update_mu ................ MonotoneMuUpdate
//   Algo::compute_direction ........ PDSearchDirCalculator + PDFullSpaceSolver + PDPerturbationHandler
//                                    (+ AugRestoSystemSolver: elimination of the n/p variables)
//   Algo::line_search .............. BacktrackingLineSearch::FindAcceptableTrialPoint (watchdog, soft
//                                    restoration phase, tiny steps) + FilterLSAcceptor (filter, switching /
//                                    Armijo conditions, second-order correction, filter resets, obj_max_inc)
//   Algo::perform_restoration ...... MinC_1NrmRestorationPhase (outer) / RestoRestorationPhase (inner)
//   RestoNlp ....................... RestoIpoptNLP,  Algo::init_resto = RestoIterateInitializer
// The restoration problem is handed to the SAME algorithm object type, as IPOPT does.
// ------------------------------------------------------------------------------------------
enum Status : int32_t {
  SOLVE_SUCCEEDED = 0, MAXITER_EXCEEDED = 1, RESTORATION_FAILED = 2, STEP_TOO_SMALL = 3, INVALID_NUMBER = 4,
  ERROR_IN_STEP_COMPUTATION = 5, INFEASIBLE_PROBLEM_DETECTED = 6,
  // results of the restoration sub-algorithm (never returned to the caller)
  RESTO_SUCCESS = 100, RESTO_CONVERGED_TO_FEASIBLE = 101, CONTINUE = 200
};

struct Options {
  int max_iter = 100; double tol = 1e-8;
  double dual_inf_tol = 1.0, constr_viol_tol = 1e-4, compl_inf_tol = 1e-4;
  double bound_relax = 1e-8, bound_push = 1e-2, bound_frac = 1e-2;
  double mu_init = 0.1, kappa_mu = 0.2, theta_mu = 1.5, kappa_eps = 10.0, tau_min = 0.99;
  double kappa_d = 1e-4, kappa_sigma = 1e10, s_max = 100.0;
  double max_grad = 100.0, scal_min = 1e-8; int scaling = 1;
  double constr_mult_init_max = 1e3;
  double dw_init = 1e-4, dw_min = 1e-20, dw_max = 1e20, dw_inc_first = 100.0, dw_inc = 8.0, dw_dec = 1.0 / 3.0;
  double gamma_theta = 1e-5, gamma_phi = 1e-8, eta_phi = 1e-8, s_theta = 1.1, s_phi = 2.3, delta = 1.0;
  double alpha_min_frac = 0.05, alpha_red = 0.5; int max_soc = 4; double kappa_soc = 0.99;
  double theta_max_fact = 1e4, theta_min_fact = 1e-4;
  double tiny_step_tol = 10 * EPS, tiny_step_y_tol = 1e-2;
  double obj_max_inc = 5.0; int max_filter_resets = 5, filter_reset_trigger = 5;
  int filter_cap = 24;      // IPOPT's filter is unbounded; entries dominated by a new one are removed, so it stays short
  int watchdog_trigger = 10, watchdog_trial_max = 3;                 // watchdog_shortened_iter_trigger, watchdog_trial_iter_max
  int max_soft_resto = 10; double soft_resto_red = 1.0 - 1e-4;      // max_soft_resto_iters, soft_resto_pderror_reduction_factor
  double slack_move = 1.8189894035458565e-12;                        // eps^(3/4)
  // restoration phase
  int resto = 1;                                                     // 0: report RESTORATION_FAILED instead of entering it (round-1 behaviour)
  double resto_rho = 1000.0, resto_eta_factor = 1.0, kappa_resto = 0.9;   // resto_penalty_parameter, resto_proximity_weight, required_infeasibility_reduction
  double bound_mult_reset_threshold = 1e3, constr_mult_reset_threshold = 0.0;
  double resto_theta_max_fact = 1e8;
  int resto_explicit = 0;      // 1: factor the restoration KKT system with the n/p variables explicit (validation of the elimination)
  // NON-REFERENCE fast mode (the scripts pass no lam_x0 / lam_g0 and leave warm_start_init_point = no): IPOPT's
  // WarmStartIterateInitializer for instances that come with a multiplier guess [3P]
  double ws_mu_init = 1e-4;                                          // mu_init a warm-started solve begins with
  double ws_bound_push = 1e-3, ws_bound_frac = 1e-3, ws_slack_push = 1e-3, ws_slack_frac = 1e-3;   // warm_start_(slack_)bound_push / frac
  double ws_mult_push = 1e-3;                                        // warm_start_mult_bound_push
  double ws_mult_init_max = 1e6;                                     // warm_start_mult_init_max
};

struct Counters {
  int n_fact = 0, n_soc_acc = 0, n_resto = 0, n_resto_iter = 0, n_watchdog = 0, n_soft = 0, n_filter_reset = 0, n_slack_adj = 0;
};

struct IterLog { double mu, f, inf_pr, inf_du, dw, alpha_pr, alpha_du; int ls; int tag; };

inline bool cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * EPS * std::fabs(bas); }

// The problem as the algorithm sees it (IpoptNLP): scaled functions, relaxed bounds.  Variables are the n0
// "structural" ones followed by `ne` pairs (n_r, p_r) of elastic variables that enter row r as + n_r - p_r and
// nowhere else (ne = 0 for the original problem, ne = m for the restoration problem).
struct Nlp {
  int n0 = 0, ne = 0, m = 0;
  int n() const { return n0 + 2 * ne; }
  std::vector<double> xL, xU, dL, dU;
  double df = 1.0;                      // objective scaling of the original problem (unscaled tolerances)
  std::vector<double> dc;               // row scaling
  virtual double eval_fg(const double* x, double mu, double* g0) = 0;
  virtual void eval_derivs(const double* x, double mu, double* grad, double* J0) = 0;     // grad [n], J0 [m][n0]
  virtual void eval_hess(double sigma, const double* y, double mu, double* W0) = 0;        // [n0][n0], at the point of eval_derivs
  virtual ~Nlp() {}
};

struct OrigNlp : Nlp {
  Instance& I; std::vector<double> ytmp;
  explicit OrigNlp(Instance& I_) : I(I_) { n0 = I.nw; ne = 0; m = I.ng; dc.assign(m, 1.0); ytmp.resize(m); }
  double eval_fg(const double* x, double, double* g0) override {
    double f = I.eval_fg(x, g0);
    for (int r = 0; r < m; ++r) g0[r] *= dc[r];
    return df * f;
  }
  void jac(double* J0) {
    I.jac_g(J0);
    for (int r = 0; r < m; ++r) if (dc[r] != 1.0) for (int c = 0; c < n0; ++c) J0[(size_t)r * n0 + c] *= dc[r];
  }
  void eval_derivs(const double* x, double, double* grad, double* J0) override {
    I.eval_derivs(x); I.grad_f(grad); jac(J0);
    for (int i = 0; i < n0; ++i) grad[i] *= df;
  }
  void eval_hess(double sigma, const double* y, double, double* W0) override {
    for (int r = 0; r < m; ++r) ytmp[r] = y[r] * dc[r];
    I.hess_l(sigma * df, ytmp.data(), W0);
  }
};

// RestoIpoptNLP:  min rho * sum(n + p) + eta(mu)/2 ||D_R (x - x_R)||^2   s.t.  d_L <= d(x) + n - p <= d_U,  x_L <= x <= x_U,  n, p >= 0
struct RestoNlp : Nlp {
  OrigNlp& O; std::vector<double> xR, dR2; double rho, eta_factor;
  RestoNlp(OrigNlp& O_, const double* x_ref, double rho_, double eta_f) : O(O_), rho(rho_), eta_factor(eta_f) {
    n0 = O.n0; ne = O.m; m = O.m; df = O.df; dc = O.dc;
    xL = O.xL; xU = O.xU; xL.resize(n(), 0.0); xU.resize(n(), INF);
    dL = O.dL; dU = O.dU;
    xR.assign(x_ref, x_ref + n0); dR2.resize(n0);
    for (int i = 0; i < n0; ++i) { double d = 1.0 / std::max(1.0, std::fabs(xR[i])); dR2[i] = d * d; }
  }
  double eta(double mu) const { return eta_factor * std::sqrt(mu); }
  double eval_fg(const double* x, double mu, double* g0) override {
    O.eval_fg(x, mu, g0);
    double q = 0, l1 = 0;
    for (int i = 0; i < n0; ++i) { double d = x[i] - xR[i]; q += dR2[i] * d * d; }
    for (int j = n0; j < n(); ++j) l1 += x[j];
    return rho * l1 + 0.5 * eta(mu) * q;
  }
  void eval_derivs(const double* x, double mu, double* grad, double* J0) override {
    O.I.eval_derivs(x); O.jac(J0);
    for (int i = 0; i < n0; ++i) grad[i] = eta(mu) * dR2[i] * (x[i] - xR[i]);
    for (int j = n0; j < n(); ++j) grad[j] = rho;
  }
  void eval_hess(double sigma, const double* y, double mu, double* W0) override {
    O.eval_hess(0.0, y, mu, W0);
    for (int i = 0; i < n0; ++i) W0[(size_t)i * n0 + i] += sigma * eta(mu) * dR2[i];
  }
};

struct Iter { std::vector<double> x, s, y, zL, zU, vL, vU; };

struct Exit { Status st; };

struct Algo {
  Nlp& P; Options o; const bool is_resto; Algo* const outer; Counters& C;
  const int n, n0, ne, m;
  Iter cur, del, dsoc, tr;
  double f = 0, f_t = 0; std::vector<double> g, g_t, grad, J0, W0, M;
  double mu = 0.1, tau = 0.99; int iter = 0;
  double resto_tol;                        // RestoConvergenceCheck may tighten the restoration problem's tolerance once
  // linear system state
  std::vector<double> Dx, Ds, Om, Sgx, Sgs, rX, rS, cvec;
  double dw_last = 0.0, dw_cur = 0.0;
  // FilterLSAcceptor
  struct FEntry { double barr, theta; };
  std::vector<FEntry> filter;
  double theta_max = -1, theta_min = -1, ref_theta = 0, ref_barr = 0, ref_gbd = 0;
  bool last_rej_filter = false; int succ_filter_rej = 0, n_filter_resets = 0;
  // BacktrackingLineSearch
  bool in_watchdog = false; int wd_short = 0, wd_trial = 0; double wd_theta = 0, wd_barr = 0, wd_gbd = 0, wd_alpha_test = 0, wd_dw = 0;
  Iter wd_iter, wd_del;
  bool in_soft = false; int soft_cnt = 0;
  bool tiny_last = false, tiny_flag = false;
  bool first_resto_iter = true;
  std::vector<IterLog>* log = nullptr;
  // info of the last iteration (log)
  double info_alpha_pr = 0, info_alpha_du = 0; int info_ls = 0, info_tag = 0;

  Algo(Nlp& P_, const Options& o_, bool resto_, Algo* outer_, Counters& C_)
      : P(P_), o(o_), is_resto(resto_), outer(outer_), C(C_), n(P_.n()), n0(P_.n0), ne(P_.ne), m(P_.m) {
    g.resize(m); g_t.resize(m); grad.resize(n); J0.resize((size_t)m * n0); W0.resize((size_t)n0 * n0);
    const int nm = (ne && o.resto_explicit) ? n : n0;
    M.resize((size_t)nm * nm);
    Dx.resize(n); Ds.resize(m); Om.resize(m); Sgx.resize(n); Sgs.resize(m); rX.resize(n); rS.resize(m); cvec.resize(m);
    resto_tol = o.tol;
  }
  static void resize(Iter& it, int n, int m) {
    it.x.assign(n, 0); it.s.assign(m, 0); it.y.assign(m, 0); it.zL.assign(n, 0); it.zU.assign(n, 0); it.vL.assign(m, 0); it.vU.assign(m, 0);
  }
  bool hasxL(int i) const { return P.xL[i] > -INF; } bool hasxU(int i) const { return P.xU[i] < INF; }
  bool hasdL(int r) const { return P.dL[r] > -INF; } bool hasdU(int r) const { return P.dU[r] < INF; }

  // ---- small helpers -----------------------------------------------------------------------
  double crow(const std::vector<double>& gg, const Iter& it, int r) const {     // residual of row r: d(x) + n - p - s
    double c = gg[r] - it.s[r];
    if (ne) c += it.x[n0 + r] - it.x[n0 + m + r];
    return c;
  }
  double theta_of(const std::vector<double>& gg, const Iter& it) const { double t = 0; for (int r = 0; r < m; ++r) t += std::fabs(crow(gg, it, r)); return t; }
  double infpr_of(const std::vector<double>& gg, const Iter& it) const { double t = 0; for (int r = 0; r < m; ++r) t = std::max(t, std::fabs(crow(gg, it, r))); return t; }
  void JtY(const double* y, double* out) const {            // J_X^T y
    for (int i = 0; i < n0; ++i) { double v = 0; for (int r = 0; r < m; ++r) v += J0[(size_t)r * n0 + i] * y[r]; out[i] = v; }
    for (int r = 0; r < ne; ++r) { out[n0 + r] = y[r]; out[n0 + m + r] = -y[r]; }
  }
  void JX(const double* dX, double* out) const {            // J_X dX
    for (int r = 0; r < m; ++r) {
      double v = 0; const double* Jr = &J0[(size_t)r * n0];
      for (int i = 0; i < n0; ++i) v += Jr[i] * dX[i];
      if (ne) v += dX[n0 + r] - dX[n0 + m + r];
      out[r] = v;
    }
  }
  // CalculateSafeSlack.  IPOPT replaces a slack that rounding has pushed to (or below) eps * min(1, mu) by
  // min(max(mu / z, s_min), slack + slack_move * max(1, |bound|)) and moves the BOUND accordingly.  Here (and in the CUDA
  // kernel) the accepted VARIABLE is moved instead (accept_trial_point), to slack = max(slack, 0) + slack_move * max(1,
  // |bound|): the multiplier-dependent alternative mu / z is dropped (it would make the repair depend on the order of
  // the primal repair and the kappa_Sigma reset of z); trial points are evaluated with their plain slacks.
  double safe_value(double sl, double /*z*/, double bnd) const { return std::max(sl, 0.0) + o.slack_move * std::max(1.0, std::fabs(bnd)); }
  bool unsafe(double sl) const { return sl < EPS * std::min(1.0, mu); }
  double slxL(const Iter& it, int i) const { return it.x[i] - P.xL[i]; }
  double slxU(const Iter& it, int i) const { return P.xU[i] - it.x[i]; }
  double sldL(const Iter& it, int r) const { return it.s[r] - P.dL[r]; }
  double sldU(const Iter& it, int r) const { return P.dU[r] - it.s[r]; }

  double barrier_of(const Iter& it, double fv, double mu_) const {
    double phi = fv;
    for (int i = 0; i < n; ++i) {
      const bool hl = hasxL(i), hu = hasxU(i);
      if (hl) { double sl = slxL(it, i); phi -= mu_ * std::log(sl); if (!hu) phi += o.kappa_d * mu_ * sl; }
      if (hu) { double sl = slxU(it, i); phi -= mu_ * std::log(sl); if (!hl) phi += o.kappa_d * mu_ * sl; }
    }
    for (int r = 0; r < m; ++r) {
      const bool hl = hasdL(r), hu = hasdU(r);
      if (hl) { double sl = sldL(it, r); phi -= mu_ * std::log(sl); if (!hu) phi += o.kappa_d * mu_ * sl; }
      if (hu) { double sl = sldU(it, r); phi -= mu_ * std::log(sl); if (!hl) phi += o.kappa_d * mu_ * sl; }
    }
    return phi;
  }
  void eval_point() {      // f, g, grad, J at cur.x
    f = P.eval_fg(cur.x.data(), mu, g.data());
    P.eval_derivs(cur.x.data(), mu, grad.data(), J0.data());
  }

  // ---- optimality error (IpoptCalculatedQuantities::curr_nlp_error / curr_barrier_error) --------
  struct Err { double du, pr, co, sd, sc, E; };
  Err error_at(const Iter& it, const std::vector<double>& gg, const std::vector<double>& gradv, double mu_) const {
    Err e{0, 0, 0, 1, 1, 0}; double sumy = 0, sumz = 0; int nz = 0;
    std::vector<double> jty(n); JtY(it.y.data(), jty.data());
    for (int i = 0; i < n; ++i) {
      e.du = std::max(e.du, std::fabs(gradv[i] + jty[i] - it.zL[i] + it.zU[i]));
      if (hasxL(i)) { e.co = std::max(e.co, std::fabs(slxL(it, i) * it.zL[i] - mu_)); sumz += std::fabs(it.zL[i]); ++nz; }
      if (hasxU(i)) { e.co = std::max(e.co, std::fabs(slxU(it, i) * it.zU[i] - mu_)); sumz += std::fabs(it.zU[i]); ++nz; }
    }
    for (int r = 0; r < m; ++r) {
      e.du = std::max(e.du, std::fabs(-it.y[r] - it.vL[r] + it.vU[r]));
      e.pr = std::max(e.pr, std::fabs(crow(gg, it, r)));
      sumy += std::fabs(it.y[r]);
      if (hasdL(r)) { e.co = std::max(e.co, std::fabs(sldL(it, r) * it.vL[r] - mu_)); sumz += std::fabs(it.vL[r]); ++nz; }
      if (hasdU(r)) { e.co = std::max(e.co, std::fabs(sldU(it, r) * it.vU[r] - mu_)); sumz += std::fabs(it.vU[r]); ++nz; }
    }
    e.sd = std::max(o.s_max, (sumy + sumz) / std::max(1, m + nz)) / o.s_max;
    e.sc = std::max(o.s_max, sumz / std::max(1, nz)) / o.s_max;
    e.E = std::max(e.du / e.sd, std::max(e.pr, e.co / e.sc));
    return e;
  }
  Err error(double mu_) const { return error_at(cur, g, grad, mu_); }
  // curr/trial_primal_dual_system_error (soft restoration phase): averaged 1-norms
  double pd_error(const Iter& it, const std::vector<double>& gg, const std::vector<double>& gradv, const std::vector<double>& jty) const {
    double du = 0, pr = 0, co = 0; int nb = 0;
    for (int i = 0; i < n; ++i) {
      du += std::fabs(gradv[i] + jty[i] - it.zL[i] + it.zU[i]);
      if (hasxL(i)) { co += std::fabs(slxL(it, i) * it.zL[i] - mu); ++nb; }
      if (hasxU(i)) { co += std::fabs(slxU(it, i) * it.zU[i] - mu); ++nb; }
    }
    for (int r = 0; r < m; ++r) {
      du += std::fabs(-it.y[r] - it.vL[r] + it.vU[r]); pr += std::fabs(crow(gg, it, r));
      if (hasdL(r)) { co += std::fabs(sldL(it, r) * it.vL[r] - mu); ++nb; }
      if (hasdU(r)) { co += std::fabs(sldU(it, r) * it.vU[r] - mu); ++nb; }
    }
    return du / (n + m) + (m > 0 ? pr / m : 0.0) + (nb > 0 ? co / nb : 0.0);
  }

  // ---- linear algebra: condensed primal-dual system -------------------------------------------------
  // (W + D_x + J^T D_s J) dX = -(r_X + J^T (D_s c + r_s));  ds = J dX + c;  dy = D_s ds + r_s.
  // Restoration problem: the n/p variables are eliminated row by row (AugRestoSystemSolver), which turns D_s into
  // Omega = 1 / (1/D_s + 1/D_n + 1/D_p) and leaves a system in the structural variables only.
  bool factor(bool use_W, const std::vector<double>& Dx_, const std::vector<double>& Ds_) {
    ++C.n_fact;
    if (&Dx_ != &Dx) Dx = Dx_;
    if (&Ds_ != &Ds) Ds = Ds_;
    const bool expl = ne && o.resto_explicit;
    const int nm = expl ? n : n0;
    std::fill(M.begin(), M.end(), 0.0);
    if (use_W) for (int a = 0; a < n0; ++a) for (int b = 0; b <= a; ++b) M[(size_t)a * nm + b] = W0[(size_t)a * n0 + b];
    for (int i = 0; i < nm; ++i) M[(size_t)i * nm + i] += Dx[i];
    for (int r = 0; r < m; ++r) {
      Om[r] = (ne && !expl) ? 1.0 / (1.0 / Ds[r] + 1.0 / Dx[n0 + r] + 1.0 / Dx[n0 + m + r]) : Ds[r];
      const double* Jr = &J0[(size_t)r * n0]; const double d = Om[r];
      int last = -1; for (int c = n0 - 1; c >= 0; --c) if (Jr[c] != 0.0) { last = c; break; }
      for (int a = 0; a <= last; ++a) { double ja = d * Jr[a]; if (ja == 0.0) continue; double* Mr = &M[(size_t)a * nm];
        for (int b = 0; b <= a; ++b) Mr[b] += ja * Jr[b]; }
      if (expl) {
        double* Mn = &M[(size_t)(n0 + r) * nm]; double* Mp = &M[(size_t)(n0 + m + r) * nm];
        for (int b = 0; b <= last; ++b) { Mn[b] += d * Jr[b]; Mp[b] -= d * Jr[b]; }
        Mn[n0 + r] += d; Mp[n0 + m + r] += d; Mp[n0 + r] -= d;
      }
    }
    return cholesky(M.data(), nm);
  }
  // Row quantities of a step whose structural part d.x[0:n0] is known (q = J0 dx).  Restoration problem: dy first,
  //     dy = Om (q + chat),  chat = c + rs/D + rp/Dp - rn/Dn,   ds = (dy - rs)/D,  dn = -(dy + rn)/Dn,  dp = (dy - rp)/Dp,
  // so that the dual equations of s, n, p hold exactly whatever the rounding of the (large) D's.
  void recover_rows(const std::vector<double>& rx, const std::vector<double>& rs, const double* cc, Iter& d) const {
    for (int r = 0; r < m; ++r) {
      double q = 0; const double* Jr = &J0[(size_t)r * n0]; for (int i = 0; i < n0; ++i) q += Jr[i] * d.x[i];
      if (ne) {
        const double Dn = Dx[n0 + r], Dp = Dx[n0 + m + r], rn = rx[n0 + r], rp = rx[n0 + m + r];
        const double chat = cc[r] + rs[r] / Ds[r] + rp / Dp - rn / Dn;
        const double dy = Om[r] * (q + chat);
        d.y[r] = dy; d.s[r] = (dy - rs[r]) / Ds[r]; d.x[n0 + r] = -(dy + rn) / Dn; d.x[n0 + m + r] = (dy - rp) / Dp;
      } else { d.s[r] = q + cc[r]; d.y[r] = Ds[r] * d.s[r] + rs[r]; }
    }
  }
  void solve_dir(const std::vector<double>& rx, const std::vector<double>& rs, const double* cc, Iter& d) {
    const bool expl = ne && o.resto_explicit;
    std::vector<double> t(m), tt(m);
    for (int r = 0; r < m; ++r) t[r] = Ds[r] * cc[r] + rs[r];
    if (ne && !expl) {
      for (int r = 0; r < m; ++r) {
        const double Dn = Dx[n0 + r], Dp = Dx[n0 + m + r];
        tt[r] = Om[r] * (cc[r] + rs[r] / Ds[r] + rx[n0 + m + r] / Dp - rx[n0 + r] / Dn);
      }
    } else tt = t;
    const int nm = expl ? n : n0;
    std::vector<double> rhs(nm);
    for (int i = 0; i < n0; ++i) { double v = rx[i]; for (int r = 0; r < m; ++r) v += J0[(size_t)r * n0 + i] * tt[r]; rhs[i] = -v; }
    if (expl) for (int r = 0; r < m; ++r) { rhs[n0 + r] = -(rx[n0 + r] + t[r]); rhs[n0 + m + r] = -(rx[n0 + m + r] - t[r]); }
    chol_solve(M.data(), nm, rhs.data());
    for (int i = 0; i < nm; ++i) d.x[i] = rhs[i];
    if (expl) {      // validation path: everything from the explicit solve
      for (int r = 0; r < m; ++r) {
        double q = 0; const double* Jr = &J0[(size_t)r * n0]; for (int i = 0; i < n0; ++i) q += Jr[i] * d.x[i];
        d.s[r] = q + d.x[n0 + r] - d.x[n0 + m + r] + cc[r]; d.y[r] = Ds[r] * d.s[r] + rs[r];
      }
    } else recover_rows(rx, rs, cc, d);
  }
  void dual_dirs(Iter& d) const {
    for (int i = 0; i < n; ++i) {
      d.zL[i] = d.zU[i] = 0;
      if (hasxL(i)) { double sl = slxL(cur, i); d.zL[i] = (mu - cur.zL[i] * d.x[i]) / sl - cur.zL[i]; }
      if (hasxU(i)) { double sl = slxU(cur, i); d.zU[i] = (mu + cur.zU[i] * d.x[i]) / sl - cur.zU[i]; }
    }
    for (int r = 0; r < m; ++r) {
      d.vL[r] = d.vU[r] = 0;
      if (hasdL(r)) { double sl = sldL(cur, r); d.vL[r] = (mu - cur.vL[r] * d.s[r]) / sl - cur.vL[r]; }
      if (hasdU(r)) { double sl = sldU(cur, r); d.vU[r] = (mu + cur.vU[r] * d.s[r]) / sl - cur.vU[r]; }
    }
  }
  double ftb_primal(const Iter& d) const {
    double a = 1.0;
    for (int i = 0; i < n; ++i) {
      if (hasxL(i) && d.x[i] < 0) a = std::min(a, -tau * slxL(cur, i) / d.x[i]);
      if (hasxU(i) && d.x[i] > 0) a = std::min(a, tau * slxU(cur, i) / d.x[i]);
    }
    for (int r = 0; r < m; ++r) {
      if (hasdL(r) && d.s[r] < 0) a = std::min(a, -tau * sldL(cur, r) / d.s[r]);
      if (hasdU(r) && d.s[r] > 0) a = std::min(a, tau * sldU(cur, r) / d.s[r]);
    }
    return a;
  }
  static double ftb_dual_of(const Iter& z, const Iter& d, double tau_, const Algo& A) {
    double a = 1.0;
    for (int i = 0; i < A.n; ++i) {
      if (d.zL[i] < 0 && A.hasxL(i)) a = std::min(a, -tau_ * z.zL[i] / d.zL[i]);
      if (d.zU[i] < 0 && A.hasxU(i)) a = std::min(a, -tau_ * z.zU[i] / d.zU[i]);
    }
    for (int r = 0; r < A.m; ++r) {
      if (d.vL[r] < 0 && A.hasdL(r)) a = std::min(a, -tau_ * z.vL[r] / d.vL[r]);
      if (d.vU[r] < 0 && A.hasdU(r)) a = std::min(a, -tau_ * z.vU[r] / d.vU[r]);
    }
    return a;
  }
  double ftb_dual(const Iter& d) const {
    if (getenv("ORACLE_DEBUG_DUAL")) {
      double a = 1.0; int wi = -1, wk = -1;
      for (int i = 0; i < n; ++i) {
        if (d.zL[i] < 0 && hasxL(i)) { double t = -tau * cur.zL[i] / d.zL[i]; if (t < a) { a = t; wi = i; wk = 0; } }
        if (d.zU[i] < 0 && hasxU(i)) { double t = -tau * cur.zU[i] / d.zU[i]; if (t < a) { a = t; wi = i; wk = 1; } }
      }
      for (int r = 0; r < m; ++r) {
        if (d.vL[r] < 0 && hasdL(r)) { double t = -tau * cur.vL[r] / d.vL[r]; if (t < a) { a = t; wi = r; wk = 2; } }
        if (d.vU[r] < 0 && hasdU(r)) { double t = -tau * cur.vU[r] / d.vU[r]; if (t < a) { a = t; wi = r; wk = 3; } }
      }
      if (wk >= 0) {
        const double z = wk == 0 ? cur.zL[wi] : wk == 1 ? cur.zU[wi] : wk == 2 ? cur.vL[wi] : cur.vU[wi];
        const double dz = wk == 0 ? d.zL[wi] : wk == 1 ? d.zU[wi] : wk == 2 ? d.vL[wi] : d.vU[wi];
        const double dp = wk < 2 ? d.x[wi] : d.s[wi];
        const double sl = wk == 0 ? slxL(cur, wi) : wk == 1 ? slxU(cur, wi) : wk == 2 ? sldL(cur, wi) : sldU(cur, wi);
        fprintf(stderr, "iter %d resto %d: a_du %g limited by kind %d idx %d (n0 %d m %d): z %g dz %g dprimal %g slack %g raw slack %g\n", iter, (int)is_resto, a, wk, wi, n0, m, z, dz, dp, sl,
                wk == 0 ? cur.x[wi] - P.xL[wi] : wk == 1 ? P.xU[wi] - cur.x[wi] : wk == 2 ? cur.s[wi] - P.dL[wi] : P.dU[wi] - cur.s[wi]);
      }
    }
    return ftb_dual_of(cur, d, tau, *this);
  }
  // gradient of the barrier function (with damping) times the step
  double grad_barr_t_delta(const Iter& d) const {
    double gbd = 0;
    for (int i = 0; i < n; ++i) {
      double gi = grad[i]; const bool hl = hasxL(i), hu = hasxU(i);
      if (hl) gi -= mu / slxL(cur, i);
      if (hu) gi += mu / slxU(cur, i);
      if (hl && !hu) gi += o.kappa_d * mu;
      if (hu && !hl) gi -= o.kappa_d * mu;
      gbd += gi * d.x[i];
    }
    for (int r = 0; r < m; ++r) {
      double gi = 0; const bool hl = hasdL(r), hu = hasdU(r);
      if (hl) gi -= mu / sldL(cur, r);
      if (hu) gi += mu / sldU(cur, r);
      if (hl && !hu) gi += o.kappa_d * mu;
      if (hu && !hl) gi -= o.kappa_d * mu;
      gbd += gi * d.s[r];
    }
    return gbd;
  }

  // ---- search direction (with inertia correction) ---------------------------------------------------
  void build_sigma_rhs() {      // Sigma's, right-hand sides and residual at the current point (current mu)
    std::vector<double> jty(n); JtY(cur.y.data(), jty.data());
    for (int i = 0; i < n; ++i) {
      double sig = 0, r_ = grad[i] + jty[i]; const bool hl = hasxL(i), hu = hasxU(i);
      if (hl) { double sl = slxL(cur, i); sig += cur.zL[i] / sl; r_ -= mu / sl; }
      if (hu) { double sl = slxU(cur, i); sig += cur.zU[i] / sl; r_ += mu / sl; }
      if (hl && !hu) r_ += o.kappa_d * mu;
      if (hu && !hl) r_ -= o.kappa_d * mu;
      Sgx[i] = sig; rX[i] = r_;
    }
    for (int r = 0; r < m; ++r) {
      double sig = 0, r_ = -cur.y[r]; const bool hl = hasdL(r), hu = hasdU(r);
      if (hl) { double sl = sldL(cur, r); sig += cur.vL[r] / sl; r_ -= mu / sl; }
      if (hu) { double sl = sldU(cur, r); sig += cur.vU[r] / sl; r_ += mu / sl; }
      if (hl && !hu) r_ += o.kappa_d * mu;
      if (hu && !hl) r_ -= o.kappa_d * mu;
      Sgs[r] = sig; rS[r] = r_;
      cvec[r] = crow(g, cur, r);
    }
  }
  void set_perturbation(double dw) {
    for (int i = 0; i < n; ++i) Dx[i] = Sgx[i] + dw;
    for (int r = 0; r < m; ++r) {
      Ds[r] = Sgs[r] + dw;
      Om[r] = (ne && !o.resto_explicit) ? 1.0 / (1.0 / Ds[r] + 1.0 / Dx[n0 + r] + 1.0 / Dx[n0 + m + r]) : Ds[r];
    }
  }
  bool compute_direction() {
    P.eval_hess(1.0, cur.y.data(), mu, W0.data());
    build_sigma_rhs();
    double dw = 0.0; bool ok = false;
    for (;;) {       // PDPerturbationHandler: delta_x = delta_s = dw, delta_c = delta_d = 0
      set_perturbation(dw);
      ok = factor(true, Dx, Ds);
      if (ok) break;
      if (dw == 0.0) dw = (dw_last == 0.0) ? o.dw_init : std::max(o.dw_min, dw_last * o.dw_dec);
      else dw = (dw_last == 0.0 || 1e5 * dw_last < dw) ? o.dw_inc_first * dw : o.dw_inc * dw;
      if (dw > o.dw_max) break;
    }
    if (!ok) return false;
    if (dw > 0) dw_last = dw;
    dw_cur = dw;
    solve_dir(rX, rS, cvec.data(), del);
    dual_dirs(del);
    return true;
  }

  // ---- FilterLSAcceptor ---------------------------------------------------------------------------------
  bool filter_acceptable(double barr, double theta) const {
    for (auto& e : filter) if (!(cmp_le(barr, e.barr, e.barr) || cmp_le(theta, e.theta, e.theta))) return false;
    return true;
  }
  void augment_filter() {
    const FEntry ne_{ref_barr - o.gamma_phi * ref_theta, (1 - o.gamma_theta) * ref_theta};
    std::vector<FEntry> keep;
    for (auto& e : filter) if (!(e.barr >= ne_.barr && e.theta >= ne_.theta)) keep.push_back(e);    // drop dominated entries
    if ((int)keep.size() >= o.filter_cap) keep.erase(keep.begin());       // bounded filter (the kernel's FILT_CAP): drop the oldest
    keep.push_back(ne_); filter.swap(keep);
  }
  void acceptor_reset() { filter.clear(); last_rej_filter = false; succ_filter_rej = 0; }
  bool is_ftype(double a_test) const {
    if (ref_theta == 0.0 && ref_gbd > 0.0 && ref_gbd < 100.0 * EPS) return true;
    return ref_gbd < 0.0 && a_test * std::pow(-ref_gbd, o.s_phi) > o.delta * std::pow(ref_theta, o.s_theta);
  }
  bool armijo(double a_test, double trial_barr) const { return cmp_le(trial_barr - ref_barr, o.eta_phi * a_test * ref_gbd, ref_barr); }
  bool acceptable_to_current_iterate(double trial_barr, double trial_theta, bool from_resto = false) const {
    if (!from_resto && trial_barr > ref_barr) {
      double basval = 1.0; if (std::fabs(ref_barr) > 10.0) basval = std::log10(std::fabs(ref_barr));
      if (std::log10(trial_barr - ref_barr) > o.obj_max_inc + basval) return false;
    }
    return cmp_le(trial_theta, (1 - o.gamma_theta) * ref_theta, ref_theta) || cmp_le(trial_barr - ref_barr, -o.gamma_phi * ref_theta, ref_barr);
  }
  bool check_acceptability(double a_test, double trial_barr, double trial_theta) {
    if (!std::isfinite(trial_barr) || !std::isfinite(trial_theta)) return false;
    if (theta_max < 0) theta_max = o.theta_max_fact * std::max(1.0, ref_theta);
    if (theta_min < 0) theta_min = o.theta_min_fact * std::max(1.0, ref_theta);
    if (theta_max > 0 && trial_theta > theta_max) return false;
    bool acc;
    if (a_test > 0.0 && is_ftype(a_test) && ref_theta <= theta_min) acc = armijo(a_test, trial_barr);
    else acc = acceptable_to_current_iterate(trial_barr, trial_theta);
    if (!acc) { last_rej_filter = false; return false; }
    acc = filter_acceptable(trial_barr, trial_theta);
    if (!acc) last_rej_filter = true;
    return acc;
  }
  double alpha_min() const {
    const double gbd = grad_barr_t_delta(del), th = theta_of(g, cur);
    double amin = o.gamma_theta;
    if (gbd < 0) {
      amin = std::min(o.gamma_theta, o.gamma_phi * th / (-gbd));
      if (th <= theta_min) amin = std::min(amin, o.delta * std::pow(th, o.s_theta) / std::pow(-gbd, o.s_phi));
    }
    return o.alpha_min_frac * amin;
  }
  void init_this_line_search() {
    if (!in_watchdog) {
      // filter reset heuristic: the last rejection was due to the filter in filter_reset_trigger successive iterations
      if (o.max_filter_resets > 0) {
        if (n_filter_resets < o.max_filter_resets) {
          if (last_rej_filter) {
            if (++succ_filter_rej >= o.filter_reset_trigger) { acceptor_reset(); ++n_filter_resets; ++C.n_filter_reset; }
          } else succ_filter_rej = 0;
        }
      }
      last_rej_filter = false;
      ref_theta = theta_of(g, cur); ref_barr = barrier_of(cur, f, mu); ref_gbd = grad_barr_t_delta(del);
    } else { ref_theta = wd_theta; ref_barr = wd_barr; ref_gbd = wd_gbd; }
  }

  // ---- trial points -----------------------------------------------------------------------------------------
  double theta_t = 0, barr_t = 0;
  bool eval_trial(double alpha, const Iter& d) {        // primal trial point; false on evaluation error (NaN / Inf)
    for (int i = 0; i < n; ++i) tr.x[i] = cur.x[i] + alpha * d.x[i];
    for (int r = 0; r < m; ++r) tr.s[r] = cur.s[r] + alpha * d.s[r];
    f_t = P.eval_fg(tr.x.data(), mu, g_t.data());
    theta_t = theta_of(g_t, tr); barr_t = barrier_of(tr, f_t, mu);
    return std::isfinite(theta_t) && std::isfinite(barr_t);
  }
  void dual_step(double a_pr, double a_du, const Iter& d) {
    for (int r = 0; r < m; ++r) { tr.y[r] = cur.y[r] + a_pr * d.y[r]; tr.vL[r] = cur.vL[r] + a_du * d.vL[r]; tr.vU[r] = cur.vU[r] + a_du * d.vU[r]; }
    for (int i = 0; i < n; ++i) { tr.zL[i] = cur.zL[i] + a_du * d.zL[i]; tr.zU[i] = cur.zU[i] + a_du * d.zU[i]; }
  }
  bool detect_tiny_step() const {
    for (int i = 0; i < n; ++i) if (std::fabs(del.x[i]) / (1 + std::fabs(cur.x[i])) > o.tiny_step_tol) return false;
    for (int r = 0; r < m; ++r) if (std::fabs(del.s[r]) / (1 + std::fabs(cur.s[r])) > o.tiny_step_tol) return false;
    return theta_of(g, cur) <= 1e-4;
  }
  void start_watchdog() {
    ++C.n_watchdog;
    in_watchdog = true; wd_iter = cur; wd_del = del; wd_trial = 0; wd_alpha_test = ftb_primal(del); wd_dw = dw_cur;
    wd_theta = ref_theta; wd_barr = ref_barr; wd_gbd = ref_gbd;
  }
  void stop_watchdog() {
    in_watchdog = false; cur = wd_iter; del = wd_del; wd_short = 0; dw_cur = wd_dw;
    eval_point();
    // The structural part dx of the step is the stored one; its row and dual parts are recomputed at the restored point
    // with the CURRENT barrier parameter (identical to the stored ones unless mu changed while the watchdog was active;
    // the CUDA kernel keeps only (dx, du) of a step and derives the rest when it needs it).
    build_sigma_rhs(); set_perturbation(wd_dw);
    recover_rows(rX, rS, cvec.data(), del);
    dual_dirs(del);
    ref_theta = wd_theta; ref_barr = wd_barr; ref_gbd = wd_gbd;
  }
  // second-order correction (FilterLSAcceptor::TrySecondOrderCorrection)
  bool try_soc(double a_test, double& alpha_primal, Iter*& actual) {
    if (o.max_soc == 0) return false;
    bool accept = false; int count = 0; double theta_old = 0, theta_trial = theta_t, a_soc = alpha_primal;
    std::vector<double> csoc(cvec);
    while (count < o.max_soc && !accept && (count == 0 || theta_trial <= o.kappa_soc * theta_old)) {
      theta_old = theta_trial;
      for (int r = 0; r < m; ++r) csoc[r] = a_soc * csoc[r] + crow(g_t, tr, r);
      solve_dir(rX, rS, csoc.data(), dsoc);
      dual_dirs(dsoc);
      a_soc = ftb_primal(dsoc);
      const bool okv = eval_trial(a_soc, dsoc);
      ++info_ls;
      if (!okv) break;
      accept = check_acceptability(a_test, barr_t, theta_t);
      if (accept) { alpha_primal = a_soc; actual = &dsoc; ++C.n_soc_acc; }
      else { ++count; theta_trial = theta_t; }
    }
    return accept;
  }
  // BacktrackingLineSearch::DoBacktrackingLineSearch
  bool do_backtracking(bool skip_first, double& alpha_primal, bool& soc_taken, int& n_steps, bool& eval_err, Iter*& actual) {
    eval_err = false; bool accept = false;
    const double a_max = ftb_primal(*actual);
    double amin = a_max;
    if (!in_watchdog) amin = alpha_min();
    alpha_primal = a_max;
    double a_test = alpha_primal;
    if (in_watchdog) a_test = wd_alpha_test;
    if (skip_first) alpha_primal *= o.alpha_red;
    while (alpha_primal > amin || n_steps == 0) {
      const bool okv = eval_trial(alpha_primal, *actual);
      ++info_ls;
      if (!in_watchdog) a_test = alpha_primal;      // IPOPT sets alpha_primal_test = alpha_primal inside the loop ...
      else a_test = wd_alpha_test;                  // ... except that the watchdog keeps testing against its reference step
      if (okv) accept = check_acceptability(a_test, barr_t, theta_t); else { accept = false; eval_err = true; }
      if (accept) break;
      if (in_watchdog) break;
      if (okv && alpha_primal == a_max && ref_theta <= theta_t) {
        accept = try_soc(a_test, alpha_primal, actual);
        if (accept) { soc_taken = true; break; }
      }
      alpha_primal *= o.alpha_red; ++n_steps;
    }
    if (accept) {     // UpdateForNextIteration: augment the filter unless the step was an Armijo-accepted f-type step
      if (!is_ftype(a_test) || !armijo(a_test, barr_t)) { augment_filter(); info_tag = soc_taken ? 'H' : 'h'; }
      else info_tag = soc_taken ? 'F' : 'f';
    } else if (in_watchdog) info_tag = 'w';
    return accept;
  }
  // BacktrackingLineSearch::TrySoftRestoStep
  bool try_soft_resto_step(const Iter& d, bool& satisfies_original) {
    satisfies_original = false;
    if (o.soft_resto_red == 0.0) return false;
    const double alpha = std::min(ftb_primal(d), ftb_dual(d));
    if (!eval_trial(alpha, d)) return false;
    dual_step(alpha, alpha, d);
    info_alpha_pr = alpha; info_alpha_du = alpha;
    if (check_acceptability(0.0, barr_t, theta_t)) { satisfies_original = true; return true; }
    std::vector<double> jty(n), grad_t(n), Jsave(J0);
    JtY(cur.y.data(), jty.data());
    const double curr_err = pd_error(cur, g, grad, jty);
    P.eval_derivs(tr.x.data(), mu, grad_t.data(), J0.data());
    JtY(tr.y.data(), jty.data());
    const double trial_err = pd_error(tr, g_t, grad_t, jty);
    J0.swap(Jsave);
    if (trial_err <= o.soft_resto_red * curr_err) { ++C.n_soft; return true; }
    P.eval_derivs(cur.x.data(), mu, grad_t.data(), Jsave.data());      // leave the model's derivative state at the current point
    return false;
  }

  // BacktrackingLineSearch::FindAcceptableTrialPoint.  On return `tr` (primal and dual) is the next iterate.
  void line_search(bool fallback) {
    info_ls = 0; info_tag = '?'; info_alpha_du = 0;
    bool goto_resto = fallback;
    init_this_line_search();
    Iter* actual = &del;
    bool accept = false, soc_taken = false; int n_steps = 0; double alpha_primal = 0.0;
    bool tiny = !goto_resto && detect_tiny_step();
    if (in_watchdog && (goto_resto || tiny)) { stop_watchdog(); goto_resto = false; tiny = false; actual = &del; }
    if (o.watchdog_trigger > 0 && !in_watchdog && !goto_resto && !tiny && !in_soft && wd_short >= o.watchdog_trigger) start_watchdog();
    if (tiny) {
      alpha_primal = ftb_primal(del);
      if (!eval_trial(alpha_primal, del)) throw Exit{INVALID_NUMBER};
      if (tiny_last) { tiny_flag = true; info_tag = 'T'; } else info_tag = 't';
      double dyn = 0; for (int r = 0; r < m; ++r) dyn = std::max(dyn, std::fabs(del.y[r]));
      tiny_last = dyn < o.tiny_step_y_tol;
      accept = true;
    } else tiny_last = false;
    if (!goto_resto && !tiny) {
      if (in_soft) {
        if (++soft_cnt > o.max_soft_resto) accept = false;
        else {
          bool sat = false;
          accept = try_soft_resto_step(*actual, sat);
          if (accept) { info_tag = 's'; if (sat) { in_soft = false; soft_cnt = 0; info_tag = 'S'; } }
        }
      } else {
        bool done = false, skip_first = false, eval_err = false;
        while (!done) {
          accept = do_backtracking(skip_first, alpha_primal, soc_taken, n_steps, eval_err, actual);
          if (in_watchdog) {
            if (accept) { in_watchdog = false; done = true; }
            else if (eval_err || ++wd_trial > o.watchdog_trial_max) { stop_watchdog(); actual = &del; skip_first = true; }
            else { done = true; accept = true; }
          } else done = true;
        }
      }
    }
    if (!accept) {
      if (!in_soft && !goto_resto) {      // try the current direction as a soft restoration step
        augment_filter();                 // PrepareRestoPhaseStart
        bool sat = false;
        accept = try_soft_resto_step(*actual, sat);
        if (accept) { if (sat) info_tag = 'S'; else { in_soft = true; info_tag = 's'; } }
      }
      if (!accept) {
        if (!in_soft) augment_filter();
        if (theta_of(g, cur) <= 1e-2 * o.tol) throw Exit{RESTORATION_FAILED};     // "Restoration phase called, but point is almost feasible"
        if (!o.resto) throw Exit{RESTORATION_FAILED};
        info_alpha_pr = alpha_primal; info_tag = 'R';
        perform_restoration();            // throws on failure; on success tr holds the new iterate
        in_soft = false; soft_cnt = 0; wd_short = 0;
      }
    } else if (!in_soft || tiny) {
      const double a_du = ftb_dual(*actual);
      dual_step(alpha_primal, a_du, *actual);
      info_alpha_pr = alpha_primal; info_alpha_du = a_du;
      if (n_steps == 0) wd_short = 0; else ++wd_short;
    }
  }

  // IpoptAlgorithm::AcceptTrialPoint: slack safeguard (bounds move), kappa_sigma correction, new current point
  void accept_trial_point() {
    const int adj0 = C.n_slack_adj;
    for (int i = 0; i < n; ++i) {
      if (hasxL(i) && unsafe(tr.x[i] - P.xL[i])) { tr.x[i] = P.xL[i] + safe_value(tr.x[i] - P.xL[i], cur.zL[i], P.xL[i]); ++C.n_slack_adj; }
      if (hasxU(i) && unsafe(P.xU[i] - tr.x[i])) { tr.x[i] = P.xU[i] - safe_value(P.xU[i] - tr.x[i], cur.zU[i], P.xU[i]); ++C.n_slack_adj; }
    }
    for (int r = 0; r < m; ++r) {
      if (hasdL(r) && unsafe(tr.s[r] - P.dL[r])) { tr.s[r] = P.dL[r] + safe_value(tr.s[r] - P.dL[r], cur.vL[r], P.dL[r]); ++C.n_slack_adj; }
      if (hasdU(r) && unsafe(P.dU[r] - tr.s[r])) { tr.s[r] = P.dU[r] - safe_value(P.dU[r] - tr.s[r], cur.vU[r], P.dU[r]); ++C.n_slack_adj; }
    }
    if (C.n_slack_adj != adj0) f_t = P.eval_fg(tr.x.data(), mu, g_t.data());      // a variable moved (by <= 1e-12 |bound|)
    auto reset = [&](double& z, double sl) { z = std::max(std::min(z, o.kappa_sigma * mu / sl), mu / (o.kappa_sigma * sl)); };
    for (int i = 0; i < n; ++i) { if (hasxL(i)) reset(tr.zL[i], slxL(tr, i)); if (hasxU(i)) reset(tr.zU[i], slxU(tr, i)); }
    for (int r = 0; r < m; ++r) { if (hasdL(r)) reset(tr.vL[r], sldL(tr, r)); if (hasdU(r)) reset(tr.vU[r], sldU(tr, r)); }
    cur = tr; f = f_t; g = g_t;
    P.eval_derivs(cur.x.data(), mu, grad.data(), J0.data());
  }

  // ---- MonotoneMuUpdate ---------------------------------------------------------------------------------
  bool mu_initialized = false;
  void update_mu() {
    const double mu_floor = std::min(o.tol, P.df * o.compl_inf_tol) / (o.kappa_eps + 1.0);
    bool tf = tiny_flag; tiny_flag = false;
    if (is_resto && !mu_initialized) { mu_initialized = true; return; }     // first restoration iteration: mu comes from the initializer
    mu_initialized = true;
    double Emu = error(mu).E;
    bool done = false;
    while ((Emu <= o.kappa_eps * mu || tf) && !done) {
      const double nm = std::max(mu_floor, std::min(o.kappa_mu * mu, std::pow(mu, o.theta_mu)));
      const bool changed = nm != mu;
      if (!changed && tf) throw Exit{STEP_TOO_SMALL};        // TINY_STEP_DETECTED: "solved to best possible numerical accuracy"
      if (!changed) break;
      mu = nm; tau = std::max(o.tau_min, 1 - mu);
      if (is_resto) { f = P.eval_fg(cur.x.data(), mu, g.data()); P.eval_derivs(cur.x.data(), mu, grad.data(), J0.data()); }   // objective depends on mu
      if (tf) { done = true; tf = false; }
      else { Emu = error(mu).E; done = !(Emu <= o.kappa_eps * mu); }
      in_soft = false; soft_cnt = 0; acceptor_reset();        // linesearch_->Reset()
    }
  }

  // ---- convergence checks -------------------------------------------------------------------------------
  Status check_convergence_orig(const double* lbg, const double* ubg) {
    const Err e = error(0.0);
    if (!std::isfinite(e.E) || !std::isfinite(f)) { if (getenv("ORACLE_DEBUG")) fprintf(stderr, "orig invalid: E %g du %g pr %g co %g f %g mu %g iter %d\n", e.E, e.du, e.pr, e.co, f, mu, iter); return INVALID_NUMBER; }
    double viol = 0;
    for (int r = 0; r < m; ++r) {
      const double gu = g[r] / P.dc[r];
      if (lbg[r] > -1e19) viol = std::max(viol, lbg[r] - gu);
      if (ubg[r] < 1e19) viol = std::max(viol, gu - ubg[r]);
    }
    if (e.E <= o.tol && e.du / P.df <= o.dual_inf_tol && viol <= o.constr_viol_tol && e.co / P.df <= o.compl_inf_tol) return SOLVE_SUCCEEDED;
    // (acceptable-level termination: the scripts set acceptable_tol = tol = 1e-8, NMPC_TT.py:260, so an "acceptable" point
    //  already satisfies the regular test and the acceptable_iter counter can never fire first)
    if (iter >= o.max_iter) return MAXITER_EXCEEDED;
    return CONTINUE;
  }
  // RestoFilterConvergenceCheck: is the restoration iterate good enough for the ORIGINAL problem's filter?
  Status check_convergence_resto() {
    Algo& A = *outer; OrigNlp& O = static_cast<RestoNlp&>(P).O;
    if (iter >= o.max_iter) return MAXITER_EXCEEDED;
    Iter ot; ot.x.assign(cur.x.begin(), cur.x.begin() + n0); ot.s = cur.s;
    std::vector<double> og(m);
    const double of = O.eval_fg(ot.x.data(), A.mu, og.data());
    const double trial_theta = A.theta_of(og, ot), trial_infpr = A.infpr_of(og, ot);
    const double curr_infpr = A.infpr_of(A.g, A.cur);
    double infpr_max = std::max(o.kappa_resto * curr_infpr, std::min(A.o.tol, A.o.constr_viol_tol));
    if (o.kappa_resto == 0.0) infpr_max = 0.0;
    Status st = CONTINUE;
    if (first_resto_iter) st = CONTINUE;                                  // always take at least one step
    else if (trial_infpr > infpr_max) st = CONTINUE;                      // not enough reduction of the original infeasibility
    else {
      // barrier function of the original problem at the restoration iterate (original mu, original bounds)
      Iter full = A.cur; full.x = ot.x; full.s = ot.s;
      const double trial_barr = A.barrier_of(full, of, A.mu);
      if (!A.filter_acceptable(trial_barr, trial_theta)) st = CONTINUE;
      else if (!A.acceptable_to_current_iterate(trial_barr, trial_theta, true)) st = CONTINUE;
      else st = RESTO_SUCCESS;
    }
    if (st == CONTINUE) {      // is the restoration problem itself solved?  then the original one is locally infeasible
      const Err e = error(0.0);
      if (!std::isfinite(e.E) || !std::isfinite(f)) { if (getenv("ORACLE_DEBUG")) fprintf(stderr, "resto invalid: E %g du %g pr %g co %g f %g mu %g iter %d\n", e.E, e.du, e.pr, e.co, f, mu, iter); return INVALID_NUMBER; }
      const double viol = infpr_of(g, cur);
      if (e.E <= resto_tol && e.du / P.df <= o.dual_inf_tol && viol <= o.constr_viol_tol && e.co / P.df <= o.compl_inf_tol) {
        if (trial_infpr <= 1e2 * resto_tol) {
          if (resto_tol > 1e-1 * A.o.tol) { resto_tol *= 1e-2; st = CONTINUE; }     // tighten once: problem only slightly infeasible
          else st = RESTO_CONVERGED_TO_FEASIBLE;
        } else st = INFEASIBLE_PROBLEM_DETECTED;
      }
    }
    first_resto_iter = false;
    return st;
  }

  // ---- restoration phase -------------------------------------------------------------------------------------
  void perform_restoration() {
    if (is_resto) {       // RestoRestorationPhase: recompute n, p for the current x, s (closed form); duals unchanged
      RestoNlp& R = static_cast<RestoNlp&>(P);
      tr = cur;
      std::vector<double> g0(m);
      R.O.eval_fg(cur.x.data(), mu, g0.data());
      for (int r = 0; r < m; ++r) {
        const double c = g0[r] - cur.s[r];
        const double a = mu / (2 * R.rho) - 0.5 * c, b = c * mu / (2 * R.rho);
        const double nv = a + std::sqrt(a * a + b);
        tr.x[n0 + r] = nv; tr.x[n0 + m + r] = c + nv;
      }
      f_t = P.eval_fg(tr.x.data(), mu, g_t.data());
      return;
    }
    ++C.n_resto;
    OrigNlp& O = static_cast<OrigNlp&>(P);
    RestoNlp R(O, cur.x.data(), o.resto_rho, o.resto_eta_factor);
    Options ro = o; ro.theta_max_fact = o.resto_theta_max_fact; ro.constr_mult_init_max = 0.0;
    Algo A(R, ro, true, this, C);
    A.log = log;
    A.iter = iter + 1;
    const Status rs = A.optimize_resto();
    C.n_resto_iter += A.iter - (iter + 1);
    // primal variables of the restoration iterate (also copied back on failure: they are what the caller gets)
    tr = cur;
    for (int i = 0; i < n0; ++i) tr.x[i] = A.cur.x[i];
    tr.s = A.cur.s;
    f_t = P.eval_fg(tr.x.data(), mu, g_t.data());
    if (rs != RESTO_SUCCESS) {
      const double orig_infpr = infpr_of(g_t, tr);
      cur.x = tr.x; cur.s = tr.s; f = f_t; g = g_t; iter = A.iter;
      Status out;
      switch (rs) {
        case STEP_TOO_SMALL: out = orig_infpr <= 1e2 * o.tol ? RESTORATION_FAILED : INFEASIBLE_PROBLEM_DETECTED; break;
        case MAXITER_EXCEEDED: out = MAXITER_EXCEEDED; break;
        case INFEASIBLE_PROBLEM_DETECTED: out = INFEASIBLE_PROBLEM_DETECTED; break;
        case RESTO_CONVERGED_TO_FEASIBLE: case RESTORATION_FAILED: out = RESTORATION_FAILED; break;
        case INVALID_NUMBER: out = INVALID_NUMBER; break;
        default: out = ERROR_IN_STEP_COMPUTATION; break;
      }
      throw Exit{out};
    }
    // bound multipliers: pretend the whole restoration phase was one primal-dual step
    Iter dz; resize(dz, n, m);
    auto bstep = [&](double z, double sl_cur, double sl_tr) { return (mu + z * (sl_cur - sl_tr)) / sl_cur - z; };
    for (int i = 0; i < n; ++i) {
      if (hasxL(i)) dz.zL[i] = bstep(cur.zL[i], slxL(cur, i), slxL(tr, i));
      if (hasxU(i)) dz.zU[i] = bstep(cur.zU[i], slxU(cur, i), slxU(tr, i));
    }
    for (int r = 0; r < m; ++r) {
      if (hasdL(r)) dz.vL[r] = bstep(cur.vL[r], sldL(cur, r), sldL(tr, r));
      if (hasdU(r)) dz.vU[r] = bstep(cur.vU[r], sldU(cur, r), sldU(tr, r));
    }
    const double a_du = ftb_dual(dz);
    double zmax = 0;
    for (int i = 0; i < n; ++i) { tr.zL[i] = cur.zL[i] + a_du * dz.zL[i]; tr.zU[i] = cur.zU[i] + a_du * dz.zU[i]; zmax = std::max(zmax, std::max(tr.zL[i], tr.zU[i])); }
    for (int r = 0; r < m; ++r) { tr.vL[r] = cur.vL[r] + a_du * dz.vL[r]; tr.vU[r] = cur.vU[r] + a_du * dz.vU[r]; zmax = std::max(zmax, std::max(tr.vL[r], tr.vU[r])); }
    if (zmax > o.bound_mult_reset_threshold) {
      for (int i = 0; i < n; ++i) { tr.zL[i] = hasxL(i) ? 1.0 : 0.0; tr.zU[i] = hasxU(i) ? 1.0 : 0.0; }
      for (int r = 0; r < m; ++r) { tr.vL[r] = hasdL(r) ? 1.0 : 0.0; tr.vU[r] = hasdU(r) ? 1.0 : 0.0; }
    }
    // constraint multipliers: constr_mult_reset_threshold = 0  =>  no least-squares estimate, y = 0
    std::fill(tr.y.begin(), tr.y.end(), 0.0);
    info_alpha_du = a_du;
    iter = A.iter - 1;
  }
  // RestoIterateInitializer
  void init_resto() {
    Algo& A = *outer; RestoNlp& R = static_cast<RestoNlp&>(P);
    resize(cur, n, m); resize(del, n, m); resize(dsoc, n, m); resize(tr, n, m);
    mu = std::max(A.mu, A.infpr_of(A.g, A.cur)); tau = std::max(o.tau_min, 1 - mu);
    for (int i = 0; i < n0; ++i) cur.x[i] = A.cur.x[i];
    cur.s = A.cur.s;
    for (int r = 0; r < m; ++r) {
      const double c = A.g[r] - A.cur.s[r];
      const double a = mu / (2 * R.rho) - 0.5 * c, b = c * mu / (2 * R.rho);
      const double nv = a + std::sqrt(a * a + b);
      cur.x[n0 + r] = nv; cur.x[n0 + m + r] = c + nv;
      cur.zL[n0 + r] = mu / nv; cur.zL[n0 + m + r] = mu / (c + nv);
    }
    for (int i = 0; i < n0; ++i) { cur.zL[i] = hasxL(i) ? std::min(R.rho, A.cur.zL[i]) : 0.0; cur.zU[i] = hasxU(i) ? std::min(R.rho, A.cur.zU[i]) : 0.0; }
    for (int r = 0; r < m; ++r) { cur.vL[r] = hasdL(r) ? std::min(R.rho, A.cur.vL[r]) : 0.0; cur.vU[r] = hasdU(r) ? std::min(R.rho, A.cur.vU[r]) : 0.0; }
    std::fill(cur.y.begin(), cur.y.end(), 0.0);        // resto.constr_mult_init_max = 0
    eval_point();
  }
  void log_iter(const Err& e) {
    if (log) log->push_back({mu, (is_resto ? f : f / P.df), e.pr, e.du, dw_cur, info_alpha_pr, info_alpha_du, info_ls, info_tag + (is_resto ? 1000 : 0)});
  }
  Status optimize_resto() {
    try {
      init_resto();
      Status st = check_convergence_resto();
      while (st == CONTINUE) {
        update_mu();
        const Err e = error(0.0);
        const bool ok = compute_direction();
        line_search(!ok);
        log_iter(e);
        accept_trial_point();
        ++iter;
        st = check_convergence_resto();
      }
      return st;
    } catch (Exit& e) { return e.st; }
  }
};

struct Result { int32_t status; int32_t iters; double f; Counters c; };

// Original problem: scaling, bounds, starting point (DefaultIterateInitializer), main loop, finalisation.
// lamx0 / lamg0 (CasADi's lam_x0 / lam_g0, unscaled; NULL or NaN in lamx0[0] = none): warm start of the multipliers, see Options::ws_*
Result solve_instance(Instance& I, const Options& o, const double* x0, const double* lbx, const double* ubx, const double* lbg,
                      const double* ubg, double* x_out, double* g_out, double* lamx_out, double* lamg_out, std::vector<IterLog>* log,
                      const double* lamx0 = nullptr, const double* lamg0 = nullptr) {
  Result res{}; Counters C;
  OrigNlp P(I);
  const int n = P.n0, m = P.m;
  Algo A(P, o, false, nullptr, C);
  A.log = log;
  Algo::resize(A.cur, n, m); Algo::resize(A.del, n, m); Algo::resize(A.dsoc, n, m); Algo::resize(A.tr, n, m);
  A.cur.x.assign(x0, x0 + n);
  // ---- gradient-based scaling at the user's starting point
  if (o.scaling) {
    P.eval_derivs(A.cur.x.data(), 0.0, A.grad.data(), A.J0.data());
    double gmax = 0; for (int i = 0; i < n; ++i) gmax = std::max(gmax, std::fabs(A.grad[i]));
    const double dfn = gmax > o.max_grad ? std::max(o.scal_min, o.max_grad / gmax) : 1.0;
    for (int r = 0; r < m; ++r) {
      double rm = 0; for (int i = 0; i < n; ++i) rm = std::max(rm, std::fabs(A.J0[(size_t)r * n + i]));
      P.dc[r] = rm > o.max_grad ? std::max(o.scal_min, o.max_grad / rm) : 1.0;
    }
    P.df = dfn;
    if (o.scaling == 2) std::fill(P.dc.begin(), P.dc.end(), 1.0);   // debug: objective scaling only
    if (o.scaling == 3) P.df = 1.0;                                  // debug: constraint scaling only
  }
  // ---- bounds (scaled, relaxed)
  P.xL.resize(n); P.xU.resize(n); P.dL.resize(m); P.dU.resize(m);
  auto relax_lo = [&](double b) { return b > -1e19 ? b - o.bound_relax * std::max(1.0, std::fabs(b)) : -INF; };
  auto relax_hi = [&](double b) { return b < 1e19 ? b + o.bound_relax * std::max(1.0, std::fabs(b)) : INF; };
  for (int i = 0; i < n; ++i) { P.xL[i] = relax_lo(lbx[i]); P.xU[i] = relax_hi(ubx[i]); }
  for (int r = 0; r < m; ++r) {
    P.dL[r] = lbg[r] > -1e19 ? relax_lo(P.dc[r] * lbg[r]) : -INF;
    P.dU[r] = ubg[r] < 1e19 ? relax_hi(P.dc[r] * ubg[r]) : INF;
  }
  const bool warm = lamx0 && lamg0 && !std::isnan(lamx0[0]);
  auto push = [&](double& v, double lo, double hi, double k1, double k2) {
    const bool hl = lo > -INF, hu = hi < INF;
    if (hl && hu) {
      const double pl = std::min(k1 * std::max(1.0, std::fabs(lo)), k2 * (hi - lo));
      const double pu = std::min(k1 * std::max(1.0, std::fabs(hi)), k2 * (hi - lo));
      v = std::max(v, lo + pl); v = std::min(v, hi - pu);
    } else if (hl) v = std::max(v, lo + k1 * std::max(1.0, std::fabs(lo)));
    else if (hu) v = std::min(v, hi - k1 * std::max(1.0, std::fabs(hi)));
  };
  Status status = MAXITER_EXCEEDED;
  try {
    // ---- starting point
    for (int i = 0; i < n; ++i) push(A.cur.x[i], P.xL[i], P.xU[i], warm ? o.ws_bound_push : o.bound_push, warm ? o.ws_bound_frac : o.bound_frac);
    A.mu = warm ? o.ws_mu_init : o.mu_init; A.tau = std::max(o.tau_min, 1 - A.mu);
    A.f = P.eval_fg(A.cur.x.data(), A.mu, A.g.data());
    for (int r = 0; r < m; ++r) { A.cur.s[r] = A.g[r]; push(A.cur.s[r], P.dL[r], P.dU[r], warm ? o.ws_slack_push : o.bound_push, warm ? o.ws_slack_frac : o.bound_frac); }
    if (warm) {   // WarmStartIterateInitializer: multipliers from the caller, pushed away from zero; v from y_d = v_U - v_L
      auto clip = [&](double v) { return std::min(std::max(v, -o.ws_mult_init_max), o.ws_mult_init_max); };
      for (int i = 0; i < n; ++i) {
        const double lz = clip(lamx0[i] * P.df);
        A.cur.zL[i] = A.hasxL(i) ? std::max(-lz, o.ws_mult_push) : 0.0; A.cur.zU[i] = A.hasxU(i) ? std::max(lz, o.ws_mult_push) : 0.0;
      }
      for (int r = 0; r < m; ++r) {
        const double y = clip(lamg0[r] * P.df / P.dc[r]);
        A.cur.y[r] = y;
        A.cur.vL[r] = A.hasdL(r) ? std::max(-y, o.ws_mult_push) : 0.0; A.cur.vU[r] = A.hasdU(r) ? std::max(y, o.ws_mult_push) : 0.0;
      }
    } else {
    for (int i = 0; i < n; ++i) { A.cur.zL[i] = A.hasxL(i) ? 1.0 : 0.0; A.cur.zU[i] = A.hasxU(i) ? 1.0 : 0.0; }
    for (int r = 0; r < m; ++r) { A.cur.vL[r] = A.hasdL(r) ? 1.0 : 0.0; A.cur.vU[r] = A.hasdU(r) ? 1.0 : 0.0; }
    }
    P.eval_derivs(A.cur.x.data(), A.mu, A.grad.data(), A.J0.data());
    if (!warm) {   // least-squares multipliers (LeastSquareMultipliers: W = 0, D_x = D_s = I)
      std::vector<double> one_n(n, 1.0), one_m(m, 1.0), rx(n), rs(m), zero(m, 0.0);
      for (int i = 0; i < n; ++i) rx[i] = A.grad[i] - A.cur.zL[i] + A.cur.zU[i];
      for (int r = 0; r < m; ++r) rs[r] = -A.cur.vL[r] + A.cur.vU[r];
      bool ok = o.constr_mult_init_max > 0 && A.factor(false, one_n, one_m);
      if (ok) {
        Iter t; Algo::resize(t, n, m);
        A.solve_dir(rx, rs, zero.data(), t);
        A.cur.y = t.y;
        double ym = 0; for (int r = 0; r < m; ++r) ym = std::max(ym, std::fabs(A.cur.y[r]));
        if (!(ym <= o.constr_mult_init_max)) std::fill(A.cur.y.begin(), A.cur.y.end(), 0.0);
      } else std::fill(A.cur.y.begin(), A.cur.y.end(), 0.0);
    }
    // ---- main loop (IpoptAlgorithm::Optimize)
    Status st = A.check_convergence_orig(lbg, ubg);
    while (st == CONTINUE) {
      A.update_mu();
      const Algo::Err e = A.error(0.0);
      const bool ok = A.compute_direction();
      A.line_search(!ok);
      A.log_iter(e);
      A.accept_trial_point();
      ++A.iter;
      st = A.check_convergence_orig(lbg, ubg);
    }
    status = st;
  } catch (Exit& e) { status = e.st; }
  // ---- finalize: honour original bounds, unscale
  for (int i = 0; i < n; ++i) x_out[i] = std::min(std::max(A.cur.x[i], lbx[i]), ubx[i]);
  std::vector<double> gu(m);
  const double fu = I.eval_fg(x_out, gu.data());   // unscaled f, g at the returned point
  if (g_out) std::memcpy(g_out, gu.data(), sizeof(double) * m);
  if (lamx_out) for (int i = 0; i < n; ++i) lamx_out[i] = (A.cur.zU[i] - A.cur.zL[i]) / P.df;
  if (lamg_out) for (int r = 0; r < m; ++r) lamg_out[r] = A.cur.y[r] * P.dc[r] / P.df;
  res.status = status; res.iters = A.iter; res.f = fu; res.c = C;
  return res;
}

// option overrides for experiments / tests (not thread safe: set before solving)
struct Override { char name[48]; double value; };
std::vector<Override> g_overrides;
void apply_overrides(Options& o) {
  for (auto& ov : g_overrides) {
    const std::string k(ov.name); const double v = ov.value;
    if (k == "resto") o.resto = (int)v; else if (k == "watchdog_trigger") o.watchdog_trigger = (int)v;
    else if (k == "max_soft_resto") o.max_soft_resto = (int)v; else if (k == "soft_resto_red") o.soft_resto_red = v;
    else if (k == "max_filter_resets") o.max_filter_resets = (int)v; else if (k == "resto_explicit") o.resto_explicit = (int)v;
    else if (k == "max_soc") o.max_soc = (int)v; else if (k == "resto_rho") o.resto_rho = v;
    else if (k == "obj_max_inc") o.obj_max_inc = v; else if (k == "mu_init") o.mu_init = v;
    else if (k == "bound_mult_reset_threshold") o.bound_mult_reset_threshold = v;
    else if (k == "ws_mu_init") o.ws_mu_init = v; else if (k == "ws_bound_push") o.ws_bound_push = v;
    else if (k == "ws_bound_frac") o.ws_bound_frac = v; else if (k == "ws_slack_push") o.ws_slack_push = v;
    else if (k == "ws_slack_frac") o.ws_slack_frac = v; else if (k == "ws_mult_push") o.ws_mult_push = v;
  }
}

}  // namespace
// ------------------------------------------------------------------------------------------
// C interface for ctypes (tests / bench only)
// ------------------------------------------------------------------------------------------
extern "C" {

struct oracle_spec { double T; int32_t N; int32_t n_obs; double w1, w2, vfov, hfov; int32_t model; };

static Spec to_spec(const oracle_spec* s) { return Spec{s->T, s->N, s->n_obs, s->w1, s->w2, s->vfov, s->hfov, s->model}; }

// function-level evaluation at (w, p): any output pointer may be NULL
int oracle_eval_traj(const oracle_spec* spec, const double* obs, const double* w, const double* p, const double* tgt,
                     double sigma, const double* lam_g,
                     double* f, double* g, double* grad, double* J, double* H, double* X);
int oracle_eval(const oracle_spec* spec, const double* obs, const double* w, const double* p,
                double sigma, const double* lam_g,
                double* f, double* g, double* grad, double* J, double* H, double* X) {
  return oracle_eval_traj(spec, obs, w, p, nullptr, sigma, lam_g, f, g, grad, J, H, X);
}
// same with a per-stage target trajectory tgt [N][2] (NULL = p[8:10] for every stage)
int oracle_eval_traj(const oracle_spec* spec, const double* obs, const double* w, const double* p, const double* tgt,
                     double sigma, const double* lam_g,
                     double* f, double* g, double* grad, double* J, double* H, double* X) {
  Spec sp = to_spec(spec); Instance I(sp, p, obs); I.tgt = tgt;
  std::vector<double> gg(I.ng);
  double fv = I.eval_fg(w, gg.data());
  if (f) *f = fv;
  if (g) std::memcpy(g, gg.data(), sizeof(double) * I.ng);
  if (X) std::memcpy(X, I.X.data(), sizeof(double) * (sp.N + 1) * NX);
  if (grad || J || H) {
    I.eval_derivs(w);
    if (grad) I.grad_f(grad);
    if (J) I.jac_g(J);
    if (H) I.hess_l(sigma, lam_g, H);
  }
  return 0;
}

// batch solve.  Instance-major arrays: p [B][11], x0 [B][nw], outputs x [B][nw], f [B], g [B][ng], ...
// bounds shared ([nw], [ng]).  obs: [n_obs][3] shared (obs_per_instance=0) or [B][n_obs][3].
// stats (optional) [B][8] : factorisations, SOC steps accepted, restoration calls, restoration iterations,
//                          watchdog starts, soft-restoration steps, filter resets, slack safeguards
int oracle_solve_traj(const oracle_spec* spec, int B, const double* p, const double* x0, const double* tgt,
                      const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                      const double* obs, int obs_per_instance, int scaling, int max_iter, double tol,
                      double* x, double* f, double* g, double* lam_x, double* lam_g,
                      int32_t* status, int32_t* iters, int32_t* stats, int nthreads);
int oracle_solve_warm(const oracle_spec* spec, int B, const double* p, const double* x0, const double* tgt,
                      const double* lam_x0, const double* lam_g0,
                      const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                      const double* obs, int obs_per_instance, int scaling, int max_iter, double tol,
                      double* x, double* f, double* g, double* lam_x, double* lam_g,
                      int32_t* status, int32_t* iters, int32_t* stats, int nthreads);
int oracle_solve(const oracle_spec* spec, int B, const double* p, const double* x0,
                 const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                 const double* obs, int obs_per_instance, int scaling, int max_iter, double tol,
                 double* x, double* f, double* g, double* lam_x, double* lam_g,
                 int32_t* status, int32_t* iters, int32_t* stats, int nthreads) {
  return oracle_solve_traj(spec, B, p, x0, nullptr, lbx, ubx, lbg, ubg, obs, obs_per_instance, scaling, max_iter, tol,
                           x, f, g, lam_x, lam_g, status, iters, stats, nthreads);
}
// same with per-instance, per-stage target trajectories tgt [B][N][2] (NULL = p[8:10] for every stage)
int oracle_solve_traj(const oracle_spec* spec, int B, const double* p, const double* x0, const double* tgt,
                      const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                      const double* obs, int obs_per_instance, int scaling, int max_iter, double tol,
                      double* x, double* f, double* g, double* lam_x, double* lam_g,
                      int32_t* status, int32_t* iters, int32_t* stats, int nthreads) {
  return oracle_solve_warm(spec, B, p, x0, tgt, nullptr, nullptr, lbx, ubx, lbg, ubg, obs, obs_per_instance, scaling, max_iter, tol,
                           x, f, g, lam_x, lam_g, status, iters, stats, nthreads);
}
// same with multiplier guesses lam_x0 [B][nw], lam_g0 [B][ng] (non-reference warm start; a NaN in lam_x0[b][0] = cold start of b)
int oracle_solve_warm(const oracle_spec* spec, int B, const double* p, const double* x0, const double* tgt,
                      const double* lam_x0, const double* lam_g0,
                      const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                      const double* obs, int obs_per_instance, int scaling, int max_iter, double tol,
                      double* x, double* f, double* g, double* lam_x, double* lam_g,
                      int32_t* status, int32_t* iters, int32_t* stats, int nthreads) {
  Spec sp = to_spec(spec); const int nw = sp.nw(), ng = sp.ng();
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      int b = next.fetch_add(1); if (b >= B) break;
      const double* ob = obs + (obs_per_instance ? (size_t)b * sp.n_obs * 3 : 0);
      Instance I(sp, p + (size_t)b * sp.npar(), ob);
      if (tgt) I.tgt = tgt + (size_t)b * 2 * sp.N;
      Options o; o.scaling = scaling; if (max_iter > 0) o.max_iter = max_iter; if (tol > 0) o.tol = tol;
      apply_overrides(o);
      std::vector<double> xo(nw);
      Result r = solve_instance(I, o, x0 + (size_t)b * nw, lbx, ubx, lbg, ubg, xo.data(),
                                g ? g + (size_t)b * ng : nullptr, lam_x ? lam_x + (size_t)b * nw : nullptr,
                                lam_g ? lam_g + (size_t)b * ng : nullptr, nullptr,
                                lam_x0 ? lam_x0 + (size_t)b * nw : nullptr, lam_g0 ? lam_g0 + (size_t)b * ng : nullptr);
      std::memcpy(x + (size_t)b * nw, xo.data(), sizeof(double) * nw);
      if (f) f[b] = r.f;
      if (status) status[b] = r.status;
      if (iters) iters[b] = r.iters;
      if (stats) {
        int32_t* s8 = stats + 8 * (size_t)b;
        s8[0] = r.c.n_fact; s8[1] = r.c.n_soc_acc; s8[2] = r.c.n_resto; s8[3] = r.c.n_resto_iter;
        s8[4] = r.c.n_watchdog; s8[5] = r.c.n_soft; s8[6] = r.c.n_filter_reset; s8[7] = r.c.n_slack_adj;
      }
    }
  };
  if (nthreads == 1) work();
  else { std::vector<std::thread> th; for (int t = 0; t < nthreads; ++t) th.emplace_back(work); for (auto& t : th) t.join(); }
  return 0;
}

// single solve with per-iteration log [max_log][9]: mu, f, inf_pr, inf_du, dw, alpha_pr, alpha_du, ls, tag (+1000 inside the restoration phase)
int oracle_solve_log(const oracle_spec* spec, const double* p, const double* x0,
                     const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                     const double* obs, int scaling, double* x, double* f, int32_t* status, int32_t* iters,
                     double* logbuf, int max_log) {
  Spec sp = to_spec(spec); Instance I(sp, p, obs);
  Options o; o.scaling = scaling; apply_overrides(o);
  std::vector<IterLog> lg;
  Result r = solve_instance(I, o, x0, lbx, ubx, lbg, ubg, x, nullptr, nullptr, nullptr, &lg);
  *f = r.f; *status = r.status; *iters = r.iters;
  for (int i = 0; i < (int)lg.size() && i < max_log; ++i) {
    double* L = logbuf + 9 * i;
    L[0] = lg[i].mu; L[1] = lg[i].f; L[2] = lg[i].inf_pr; L[3] = lg[i].inf_du; L[4] = lg[i].dw;
    L[5] = lg[i].alpha_pr; L[6] = lg[i].alpha_du; L[7] = lg[i].ls; L[8] = lg[i].tag;
  }
  return (int)lg.size();
}

// option overrides for experiments and tests (name as in Options; not thread safe).  oracle_clear_options() removes them.
int oracle_set_option(const char* name, double value) {
  Override ov{}; std::strncpy(ov.name, name, sizeof(ov.name) - 1); ov.value = value;
  for (auto& e : g_overrides) if (std::strcmp(e.name, ov.name) == 0) { e.value = value; return 0; }
  g_overrides.push_back(ov);
  return 0;
}
int oracle_clear_options(void) { g_overrides.clear(); return 0; }

}  // extern "C"
