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
//   Not restated: IPOPT's restoration phase and watchdog (an instance that would enter
//   restoration is returned with status RESTORATION_NEEDED and counted as not converged).
//
// Build:  g++ -O3 -march=native -shared -fPIC -o oracle/_build/libnmpc_oracle.so oracle/nmpc_oracle.cpp -lpthread
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
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
struct Spec {
  double T; int N; int n_obs; double w1, w2, vfov, hfov;
  int rows() const { return 5 + n_obs; }
  int nw() const { return NU * N; }
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
  const double* p;      // 11
  const double* obs;    // n_obs x 3 : cx, cy, r_uav + r_obs
  int nw, ng, rows;
  // work (valid after eval_point)
  std::vector<double> X;        // (N+1) x 8
  std::vector<double> Sx;       // (N+1) x 8 x nw  sensitivities dX_k/dw
  std::vector<J6> cost;         // N
  std::vector<Jet<2>> obsj;     // (N+1) x n_obs
  std::vector<Jet<3>> dyn;      // N x 3   (theta, psi, v) -> T*rhs rows 0..2

  // optional per-stage predicted target [N][2] (the reference keeps (x_t, y_t) = p[8:10] over the horizon)
  const double* tgt = nullptr;
  double xt(int k) const { return tgt ? tgt[2 * k] : p[8]; }
  double yt(int k) const { return tgt ? tgt[2 * k + 1] : p[9]; }

  Instance(const Spec& s, const double* p_, const double* obs_) : sp(s), p(p_), obs(obs_) {
    nw = sp.nw(); ng = sp.ng(); rows = sp.rows();
    X.resize((sp.N + 1) * NX);
  }

  void rollout(const double* w) {
    for (int i = 0; i < NX; ++i) X[i] = p[i];
    for (int k = 0; k < sp.N; ++k) {
      const double* st = &X[k * NX]; const double* u = w + NU * k; double* nx = &X[(k + 1) * NX];
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
      f += stage_cost_literal<D>(sp, st[0], st[1], st[2], st[5], st[6], st[7], xt(k), yt(k), dtan, dsin, dcos, dsqrt, dsq).v;
    }
    for (int k = 0; k <= sp.N; ++k) {
      const double* st = &X[k * NX]; double* gk = g + k * rows;
      gk[0] = st[2]; gk[1] = st[3]; gk[2] = st[5]; gk[3] = st[6]; gk[4] = st[7];
      for (int j = 0; j < sp.n_obs; ++j) {
        double dx = st[0] - obs[3 * j], dy = st[1] - obs[3 * j + 1];
        gk[5 + j] = -std::sqrt(dx * dx + dy * dy) + obs[3 * j + 2];
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
      cost[k] = stage_cost_literal<J6>(sp, v[0], v[1], v[2], v[3], v[4], v[5], xt(k), yt(k), j6tan, j6sin, j6cos, j6sqrt, j6sq);
      Jet<3> th = Jet<3>::var(st[3], 0), ps = Jet<3>::var(st[4], 1), vv = Jet<3>::var(w[NU * k], 2);
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
      for (int c = 0; c < NU * k; ++c) {   // only columns of earlier stages are non-zero
        for (int r = 0; r < NX; ++r) S1[r * nw + c] = S0[r * nw + c];
        for (int r = 0; r < 3; ++r)
          S1[r * nw + c] += dyn[k * 3 + r].g[0] * S0[3 * nw + c] + dyn[k * 3 + r].g[1] * S0[4 * nw + c];
      }
      int c0 = NU * k;
      for (int r = 0; r < 3; ++r) S1[r * nw + c0] = dyn[k * 3 + r].g[2];
      for (int i = 1; i < NU; ++i) S1[(2 + i) * nw + c0 + i] = sp.T;   // rows 3..7 <- controls 1..5
    }
  }
  void grad_f(double* grad) const {
    std::fill(grad, grad + nw, 0.0);
    for (int k = 1; k < sp.N; ++k) {   // stage 0 is constant in w
      const double* S = &Sx[(size_t)k * NX * nw];
      for (int i = 0; i < 6; ++i) {
        double gi = cost[k].g[i]; const double* row = S + COST_IDX[i] * nw;
        for (int c = 0; c < NU * k; ++c) grad[c] += gi * row[c];
      }
    }
  }
  // J : ng x nw, row-major
  void jac_g(double* J) const {
    std::fill(J, J + (size_t)ng * nw, 0.0);
    static const int LIN[5] = {2, 3, 5, 6, 7};
    for (int k = 1; k <= sp.N; ++k) {
      const double* S = &Sx[(size_t)k * NX * nw];
      for (int i = 0; i < 5; ++i) {
        double* Jr = J + (size_t)(k * rows + i) * nw; const double* row = S + LIN[i] * nw;
        for (int c = 0; c < NU * k; ++c) Jr[c] = row[c];
      }
      for (int j = 0; j < sp.n_obs; ++j) {
        double* Jr = J + (size_t)(k * rows + 5 + j) * nw; const Jet<2>& o = obsj[k * sp.n_obs + j];
        for (int c = 0; c < NU * k; ++c) Jr[c] = o.g[0] * S[c] + o.g[1] * S[nw + c];
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
      for (int i = 0; i < 5; ++i) l[LIN[i]] += lam[k * rows + i];
      for (int j = 0; j < sp.n_obs; ++j) {
        l[0] += lam[k * rows + 5 + j] * obsj[k * sp.n_obs + j].g[0];
        l[1] += lam[k * rows + 5 + j] * obsj[k * sp.n_obs + j].g[1];
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
          const Jet<2>& o = obsj[k * sp.n_obs + j]; double l = lam[k * rows + 5 + j];
          for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) H[a][b] += l * o.h[a][b];
        }
      if (k < N) {   // dynamics curvature weighted by the next-stage adjoint; vars (theta=3, psi=4, v=8)
        static const int DI[3] = {3, 4, 8};
        const double* ln = &lamx[(k + 1) * NX];
        for (int r = 0; r < 3; ++r)
          for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H[DI[a]][DI[b]] += ln[r] * dyn[k * 3 + r].h[a][b];
      }
      int ncol = std::min(nw, NU * (k + 1));   // Z's non-zero columns: controls of stages <= k
      std::fill(Z.begin(), Z.end(), 0.0);
      const double* S = &Sx[(size_t)k * NX * nw];
      for (int r = 0; r < NX; ++r) for (int c = 0; c < NU * k; ++c) Z[(size_t)r * nw + c] = S[(size_t)r * nw + c];
      if (k < N) for (int i = 0; i < NU; ++i) Z[(size_t)(NX + i) * nw + NU * k + i] = 1.0;
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
// ------------------------------------------------------------------------------------------
enum Status : int32_t {
  SOLVE_SUCCEEDED = 0, MAXITER_EXCEEDED = 1, RESTORATION_NEEDED = 2, STEP_TOO_SMALL = 3, INVALID_NUMBER = 4,
  PERTURBATION_FAILED = 5
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
  double tiny_step_tol = 10 * EPS;
};

struct IterLog { double mu, f, inf_pr, inf_du, dw, alpha_pr, alpha_du; int ls; };

struct Result { int32_t status; int32_t iters; double f; int n_fact; int n_soc_acc; };

inline bool cmp_le(double lhs, double rhs, double bas) { return lhs - rhs <= 10.0 * EPS * std::fabs(bas); }

struct Ipm {
  Instance& P; Options o; int n, m;
  std::vector<double> xL, xU, dL, dU;            // relaxed (scaled for d) bounds; +-INF when absent
  std::vector<double> dc; double df = 1.0;       // scaling
  std::vector<double> x, s, y, zL, zU, vL, vU;
  std::vector<double> g, grad, J, W, M, rhs, c;
  std::vector<double> dx, ds, dy, dzL, dzU, dvL, dvU;
  std::vector<std::pair<double, double>> filter;   // (theta, phi) margins
  double mu, tau, dw_last = 0.0, theta_max = 0, theta_min = 0;
  std::vector<IterLog>* log = nullptr;
  int n_fact = 0, n_soc_acc = 0;

  Ipm(Instance& P_, const Options& o_) : P(P_), o(o_), n(P_.nw), m(P_.ng) {}

  double feval(const double* xx, double* gg) {   // scaled f, g
    double f = P.eval_fg(xx, gg);
    for (int i = 0; i < m; ++i) gg[i] *= dc[i];
    return df * f;
  }
  void deriv(const double* xx) {   // scaled grad, J at xx
    P.eval_derivs(xx); P.grad_f(grad.data()); P.jac_g(J.data());
    for (int i = 0; i < n; ++i) grad[i] *= df;
    for (int r = 0; r < m; ++r) if (dc[r] != 1.0) for (int cidx = 0; cidx < n; ++cidx) J[(size_t)r * n + cidx] *= dc[r];
  }
  static void push(double& v, double lo, double hi, double k1, double k2) {
    bool hl = lo > -INF, hu = hi < INF;
    if (hl && hu) {
      double pl = std::min(k1 * std::max(1.0, std::fabs(lo)), k2 * (hi - lo));
      double pu = std::min(k1 * std::max(1.0, std::fabs(hi)), k2 * (hi - lo));
      v = std::max(v, lo + pl); v = std::min(v, hi - pu);
    } else if (hl) v = std::max(v, lo + k1 * std::max(1.0, std::fabs(lo)));
    else if (hu) v = std::min(v, hi - k1 * std::max(1.0, std::fabs(hi)));
  }
  // error measure E_mu (scaled)
  double error(double mu_, double* o_du = nullptr, double* o_pr = nullptr, double* o_co = nullptr) {
    double du = 0, pr = 0, co = 0, sumy = 0, sumz = 0; int nz = 0;
    for (int i = 0; i < n; ++i) {
      double gl = grad[i];
      for (int r = 0; r < m; ++r) gl += J[(size_t)r * n + i] * y[r];
      gl += -zL[i] + zU[i];
      du = std::max(du, std::fabs(gl));
      if (xL[i] > -INF) { co = std::max(co, std::fabs((x[i] - xL[i]) * zL[i] - mu_)); sumz += std::fabs(zL[i]); ++nz; }
      if (xU[i] < INF) { co = std::max(co, std::fabs((xU[i] - x[i]) * zU[i] - mu_)); sumz += std::fabs(zU[i]); ++nz; }
    }
    for (int r = 0; r < m; ++r) {
      du = std::max(du, std::fabs(-y[r] - vL[r] + vU[r]));
      pr = std::max(pr, std::fabs(g[r] - s[r]));
      sumy += std::fabs(y[r]);
      if (dL[r] > -INF) { co = std::max(co, std::fabs((s[r] - dL[r]) * vL[r] - mu_)); sumz += std::fabs(vL[r]); ++nz; }
      if (dU[r] < INF) { co = std::max(co, std::fabs((dU[r] - s[r]) * vU[r] - mu_)); sumz += std::fabs(vU[r]); ++nz; }
    }
    double sd = std::max(o.s_max, (sumy + sumz) / std::max(1, m + nz)) / o.s_max;
    double sc = std::max(o.s_max, sumz / std::max(1, nz)) / o.s_max;
    if (o_du) *o_du = du; if (o_pr) *o_pr = pr; if (o_co) *o_co = co;
    return std::max(du / sd, std::max(pr, co / sc));
  }
  double barrier(const double* xx, const double* ss, double f) const {
    double phi = f;
    for (int i = 0; i < n; ++i) {
      bool hl = xL[i] > -INF, hu = xU[i] < INF;
      if (hl) phi -= mu * std::log(xx[i] - xL[i]);
      if (hu) phi -= mu * std::log(xU[i] - xx[i]);
      if (hl && !hu) phi += o.kappa_d * mu * (xx[i] - xL[i]);
      if (hu && !hl) phi += o.kappa_d * mu * (xU[i] - xx[i]);
    }
    for (int r = 0; r < m; ++r) {
      bool hl = dL[r] > -INF, hu = dU[r] < INF;
      if (hl) phi -= mu * std::log(ss[r] - dL[r]);
      if (hu) phi -= mu * std::log(dU[r] - ss[r]);
      if (hl && !hu) phi += o.kappa_d * mu * (ss[r] - dL[r]);
      if (hu && !hl) phi += o.kappa_d * mu * (dU[r] - ss[r]);
    }
    return phi;
  }
  // build condensed matrix for given (use_W, Dx, Ds) into M and factor.  returns false if not PD
  bool factor(bool use_W, const std::vector<double>& Dx, const std::vector<double>& Ds) {
    ++n_fact;
    if (use_W) M = W; else std::fill(M.begin(), M.end(), 0.0);
    for (int i = 0; i < n; ++i) M[(size_t)i * n + i] += Dx[i];
    for (int r = 0; r < m; ++r) {
      const double* Jr = &J[(size_t)r * n]; double d = Ds[r];
      int last = -1; for (int cidx = n - 1; cidx >= 0; --cidx) if (Jr[cidx] != 0.0) { last = cidx; break; }
      for (int a = 0; a <= last; ++a) { double ja = d * Jr[a]; if (ja == 0.0) continue; double* Mr = &M[(size_t)a * n];
        for (int b = 0; b <= a; ++b) Mr[b] += ja * Jr[b]; }
    }
    return cholesky(M.data(), n);   // only the lower triangle is referenced
  }
  // solve the reduced primal-dual system for constraint residual cc (c or c_soc); M must be factored with (Dx, Ds)
  void solve_dir(const std::vector<double>& Ds, const std::vector<double>& rx, const std::vector<double>& rs,
                 const double* cc, double* ddx, double* dds, double* ddy) {
    std::vector<double> t(m);
    for (int r = 0; r < m; ++r) t[r] = Ds[r] * cc[r] + rs[r];
    for (int i = 0; i < n; ++i) { double v = rx[i]; for (int r = 0; r < m; ++r) v += J[(size_t)r * n + i] * t[r]; ddx[i] = -v; }
    chol_solve(M.data(), n, ddx);
    for (int r = 0; r < m; ++r) {
      double jd = 0; for (int i = 0; i < n; ++i) jd += J[(size_t)r * n + i] * ddx[i];
      dds[r] = jd + cc[r]; ddy[r] = Ds[r] * dds[r] + rs[r];
    }
  }
  void dual_dirs(const double* ddx, const double* dds) {
    for (int i = 0; i < n; ++i) {
      dzL[i] = dzU[i] = 0;
      if (xL[i] > -INF) { double sl = x[i] - xL[i]; dzL[i] = (mu - zL[i] * ddx[i]) / sl - zL[i]; }
      if (xU[i] < INF) { double sl = xU[i] - x[i]; dzU[i] = (mu + zU[i] * ddx[i]) / sl - zU[i]; }
    }
    for (int r = 0; r < m; ++r) {
      dvL[r] = dvU[r] = 0;
      if (dL[r] > -INF) { double sl = s[r] - dL[r]; dvL[r] = (mu - vL[r] * dds[r]) / sl - vL[r]; }
      if (dU[r] < INF) { double sl = dU[r] - s[r]; dvU[r] = (mu + vU[r] * dds[r]) / sl - vU[r]; }
    }
  }
  double ftb_primal(const double* ddx, const double* dds) const {
    double a = 1.0;
    for (int i = 0; i < n; ++i) {
      if (xL[i] > -INF && ddx[i] < 0) a = std::min(a, -tau * (x[i] - xL[i]) / ddx[i]);
      if (xU[i] < INF && ddx[i] > 0) a = std::min(a, tau * (xU[i] - x[i]) / ddx[i]);
    }
    for (int r = 0; r < m; ++r) {
      if (dL[r] > -INF && dds[r] < 0) a = std::min(a, -tau * (s[r] - dL[r]) / dds[r]);
      if (dU[r] < INF && dds[r] > 0) a = std::min(a, tau * (dU[r] - s[r]) / dds[r]);
    }
    return a;
  }
  double ftb_dual() const {
    double a = 1.0;
    for (int i = 0; i < n; ++i) {
      if (dzL[i] < 0 && xL[i] > -INF) a = std::min(a, -tau * zL[i] / dzL[i]);
      if (dzU[i] < 0 && xU[i] < INF) a = std::min(a, -tau * zU[i] / dzU[i]);
    }
    for (int r = 0; r < m; ++r) {
      if (dvL[r] < 0 && dL[r] > -INF) a = std::min(a, -tau * vL[r] / dvL[r]);
      if (dvU[r] < 0 && dU[r] < INF) a = std::min(a, -tau * vU[r] / dvU[r]);
    }
    return a;
  }
  bool filter_ok(double th, double ph) const {
    for (auto& e : filter) if (!(th < e.first || ph < e.second)) return false;
    return true;
  }

  Result solve(const double* x0, const double* lbx, const double* ubx, const double* lbg, const double* ubg,
               double* x_out, double* g_out, double* lamx_out, double* lamg_out) {
    Result res{};
    x.assign(x0, x0 + n); s.assign(m, 0); y.assign(m, 0);
    g.resize(m); grad.resize(n); J.resize((size_t)m * n); W.resize((size_t)n * n); M.resize((size_t)n * n);
    c.resize(m); dx.resize(n); ds.resize(m); dy.resize(m); dzL.resize(n); dzU.resize(n); dvL.resize(m); dvU.resize(m);
    dc.assign(m, 1.0); df = 1.0;
    // ---- gradient-based scaling at the user's starting point
    if (o.scaling) {
      deriv(x.data());
      double gmax = 0; for (int i = 0; i < n; ++i) gmax = std::max(gmax, std::fabs(grad[i]));
      double dfn = gmax > o.max_grad ? std::max(o.scal_min, o.max_grad / gmax) : 1.0;
      for (int r = 0; r < m; ++r) {
        double rm = 0; for (int i = 0; i < n; ++i) rm = std::max(rm, std::fabs(J[(size_t)r * n + i]));
        dc[r] = rm > o.max_grad ? std::max(o.scal_min, o.max_grad / rm) : 1.0;
      }
      df = dfn;
      if (o.scaling == 2) std::fill(dc.begin(), dc.end(), 1.0);   // debug: objective scaling only
      if (o.scaling == 3) df = 1.0;                                // debug: constraint scaling only
    }
    // ---- bounds (scaled, relaxed)
    xL.resize(n); xU.resize(n); dL.resize(m); dU.resize(m);
    auto relax_lo = [&](double b) { return b > -1e19 ? b - o.bound_relax * std::max(1.0, std::fabs(b)) : -INF; };
    auto relax_hi = [&](double b) { return b < 1e19 ? b + o.bound_relax * std::max(1.0, std::fabs(b)) : INF; };
    for (int i = 0; i < n; ++i) { xL[i] = relax_lo(lbx[i]); xU[i] = relax_hi(ubx[i]); }
    for (int r = 0; r < m; ++r) {
      dL[r] = lbg[r] > -1e19 ? relax_lo(dc[r] * lbg[r]) : -INF;
      dU[r] = ubg[r] < 1e19 ? relax_hi(dc[r] * ubg[r]) : INF;
    }
    // ---- starting point
    for (int i = 0; i < n; ++i) push(x[i], xL[i], xU[i], o.bound_push, o.bound_frac);
    double f = feval(x.data(), g.data());
    for (int r = 0; r < m; ++r) { s[r] = g[r]; push(s[r], dL[r], dU[r], o.bound_push, o.bound_frac); }
    zL.assign(n, 0); zU.assign(n, 0); vL.assign(m, 0); vU.assign(m, 0);
    for (int i = 0; i < n; ++i) { if (xL[i] > -INF) zL[i] = 1; if (xU[i] < INF) zU[i] = 1; }
    for (int r = 0; r < m; ++r) { if (dL[r] > -INF) vL[r] = 1; if (dU[r] < INF) vU[r] = 1; }
    deriv(x.data());
    {   // least-squares multipliers
      std::vector<double> one_n(n, 1.0), one_m(m, 1.0), rx(n), rs(m), zero(m, 0.0), tx(n), ts(m);
      for (int i = 0; i < n; ++i) rx[i] = grad[i] - zL[i] + zU[i];
      for (int r = 0; r < m; ++r) rs[r] = -vL[r] + vU[r];
      bool ok = factor(false, one_n, one_m);
      if (ok) {
        solve_dir(one_m, rx, rs, zero.data(), tx.data(), ts.data(), y.data());
        double ym = 0; for (int r = 0; r < m; ++r) ym = std::max(ym, std::fabs(y[r]));
        if (!(ym <= o.constr_mult_init_max)) std::fill(y.begin(), y.end(), 0.0);
      } else std::fill(y.begin(), y.end(), 0.0);
    }
    mu = o.mu_init; tau = std::max(o.tau_min, 1 - mu);
    double mu_floor = std::min(o.tol, o.compl_inf_tol) / (o.kappa_eps + 1.0);
    {
      double th0 = 0; for (int r = 0; r < m; ++r) th0 += std::fabs(g[r] - s[r]);
      theta_max = o.theta_max_fact * std::max(1.0, th0); theta_min = o.theta_min_fact * std::max(1.0, th0);
    }
    filter.clear();
    std::vector<double> Dx(n), Ds(m), rx(n), rs(m), xt(n), st(m), gt(m), csoc(m), dx2(n), ds2(m), dy2(m);
    int iter = 0; int tiny_count = 0; bool tiny_flag = false;
    Status status = MAXITER_EXCEEDED;
    for (;;) {
      // -------- convergence check
      double du, pr, co;
      double E0 = error(0.0, &du, &pr, &co);
      if (!std::isfinite(E0) || !std::isfinite(f)) { status = INVALID_NUMBER; break; }
      {
        // unscaled quantities for the absolute tolerances
        double viol = 0;
        for (int r = 0; r < m; ++r) {
          double gu = g[r] / dc[r];
          if (lbg[r] > -1e19) viol = std::max(viol, lbg[r] - gu);
          if (ubg[r] < 1e19) viol = std::max(viol, gu - ubg[r]);
        }
        if (E0 <= o.tol && du / df <= o.dual_inf_tol && viol <= o.constr_viol_tol && co / df <= o.compl_inf_tol) {
          status = SOLVE_SUCCEEDED; break;
        }
      }
      if (iter >= o.max_iter) { status = MAXITER_EXCEEDED; break; }
      // -------- barrier parameter update
      {
        double Emu = error(mu);
        while ((Emu <= o.kappa_eps * mu || tiny_flag) && mu > mu_floor) {
          double nm = std::max(mu_floor, std::min(o.kappa_mu * mu, std::pow(mu, o.theta_mu)));
          mu = nm; tau = std::max(o.tau_min, 1 - mu); filter.clear(); tiny_flag = false;
          Emu = error(mu);
        }
        if (tiny_flag && mu <= mu_floor) { status = STEP_TOO_SMALL; break; }
      }
      // -------- search direction
      P.hess_l(df, [&] { for (int r = 0; r < m; ++r) c[r] = y[r] * dc[r]; return c.data(); }(), W.data());
      for (int r = 0; r < m; ++r) c[r] = g[r] - s[r];
      double theta = 0; for (int r = 0; r < m; ++r) theta += std::fabs(c[r]);
      std::vector<double> Sx_(n), Ss_(m);
      for (int i = 0; i < n; ++i) {
        double sig = 0, r_ = grad[i];
        for (int r = 0; r < m; ++r) r_ += J[(size_t)r * n + i] * y[r];
        bool hl = xL[i] > -INF, hu = xU[i] < INF;
        if (hl) { sig += zL[i] / (x[i] - xL[i]); r_ -= mu / (x[i] - xL[i]); }
        if (hu) { sig += zU[i] / (xU[i] - x[i]); r_ += mu / (xU[i] - x[i]); }
        if (hl && !hu) r_ += o.kappa_d * mu; if (hu && !hl) r_ -= o.kappa_d * mu;
        Sx_[i] = sig; rx[i] = r_;
      }
      for (int r = 0; r < m; ++r) {
        double sig = 0, r_ = -y[r];
        bool hl = dL[r] > -INF, hu = dU[r] < INF;
        if (hl) { sig += vL[r] / (s[r] - dL[r]); r_ -= mu / (s[r] - dL[r]); }
        if (hu) { sig += vU[r] / (dU[r] - s[r]); r_ += mu / (dU[r] - s[r]); }
        if (hl && !hu) r_ += o.kappa_d * mu; if (hu && !hl) r_ -= o.kappa_d * mu;
        Ss_[r] = sig; rs[r] = r_;
      }
      double dw = 0.0; bool ok = false;
      for (;;) {
        for (int i = 0; i < n; ++i) Dx[i] = Sx_[i] + dw;
        for (int r = 0; r < m; ++r) Ds[r] = Ss_[r] + dw;
        ok = factor(true, Dx, Ds);
        if (ok) break;
        if (dw == 0.0) dw = (dw_last == 0.0) ? o.dw_init : std::max(o.dw_min, dw_last * o.dw_dec);
        else dw = (dw_last == 0.0 || 1e5 * dw_last < dw) ? o.dw_inc_first * dw : o.dw_inc * dw;
        if (dw > o.dw_max) break;
      }
      if (!ok) { status = PERTURBATION_FAILED; break; }
      if (dw > 0) dw_last = dw;
      solve_dir(Ds, rx, rs, c.data(), dx.data(), ds.data(), dy.data());
      dual_dirs(dx.data(), ds.data());
      double a_pr_max = ftb_primal(dx.data(), ds.data());
      double a_du = ftb_dual();
      // -------- line search
      double phi = barrier(x.data(), s.data(), f);
      double gbd = 0;   // directional derivative of the barrier function
      for (int i = 0; i < n; ++i) {
        double gi = grad[i]; bool hl = xL[i] > -INF, hu = xU[i] < INF;
        if (hl) gi -= mu / (x[i] - xL[i]); if (hu) gi += mu / (xU[i] - x[i]);
        if (hl && !hu) gi += o.kappa_d * mu; if (hu && !hl) gi -= o.kappa_d * mu;
        gbd += gi * dx[i];
      }
      for (int r = 0; r < m; ++r) {
        double gi = 0; bool hl = dL[r] > -INF, hu = dU[r] < INF;
        if (hl) gi -= mu / (s[r] - dL[r]); if (hu) gi += mu / (dU[r] - s[r]);
        if (hl && !hu) gi += o.kappa_d * mu; if (hu && !hl) gi -= o.kappa_d * mu;
        gbd += gi * ds[r];
      }
      // tiny step?
      bool tiny = true;
      for (int i = 0; i < n && tiny; ++i) if (std::fabs(dx[i]) / (1 + std::fabs(x[i])) > o.tiny_step_tol) tiny = false;
      for (int r = 0; r < m && tiny; ++r) if (std::fabs(ds[r]) / (1 + std::fabs(s[r])) > o.tiny_step_tol) tiny = false;
      if (tiny && theta > 1e-4) tiny = false;
      double alpha = a_pr_max; bool accepted = false; int ls = 0; double f_t = f;
      const double* use_dx = dx.data(); const double* use_ds = ds.data(); const double* use_dy = dy.data();
      double alpha_test = alpha;   // alpha used in the Armijo / switching tests
      auto is_ftype = [&](double a) { return gbd < 0 && a * std::pow(-gbd, o.s_phi) > o.delta * std::pow(theta, o.s_theta); };
      auto armijo = [&](double a, double ph_t) { return cmp_le(ph_t - phi, o.eta_phi * a * gbd, phi); };
      auto acceptable = [&](double a_test, double th_t, double ph_t) {
        if (!std::isfinite(th_t) || !std::isfinite(ph_t)) return false;
        if (th_t > theta_max) return false;
        bool acc;
        if (a_test > 0 && is_ftype(a_test) && theta <= theta_min) acc = armijo(a_test, ph_t);
        else acc = cmp_le(th_t, (1 - o.gamma_theta) * theta, theta) || cmp_le(ph_t - phi, -o.gamma_phi * theta, phi);
        if (!acc) return false;
        return filter_ok(th_t, ph_t);
      };
      if (tiny) {
        for (int i = 0; i < n; ++i) xt[i] = x[i] + alpha * dx[i];
        for (int r = 0; r < m; ++r) st[r] = s[r] + alpha * ds[r];
        f_t = feval(xt.data(), gt.data()); accepted = true; ++tiny_count; tiny_flag = true;
        if (tiny_count >= 2 && mu <= mu_floor) { status = STEP_TOO_SMALL; }
      } else {
        tiny_count = 0;
        double amin = o.gamma_theta;
        if (gbd < 0) {
          amin = std::min(o.gamma_theta, o.gamma_phi * theta / (-gbd));
          if (theta <= theta_min) amin = std::min(amin, o.delta * std::pow(theta, o.s_theta) / std::pow(-gbd, o.s_phi));
        }
        amin *= o.alpha_min_frac;
        bool first = true;
        while (alpha > amin || first) {
          ++ls;
          for (int i = 0; i < n; ++i) xt[i] = x[i] + alpha * dx[i];
          for (int r = 0; r < m; ++r) st[r] = s[r] + alpha * ds[r];
          f_t = feval(xt.data(), gt.data());
          double th_t = 0; for (int r = 0; r < m; ++r) th_t += std::fabs(gt[r] - st[r]);
          double ph_t = barrier(xt.data(), st.data(), f_t);
          alpha_test = alpha;
          if (acceptable(alpha, th_t, ph_t)) { accepted = true; break; }
          if (first && o.max_soc > 0 && th_t >= theta && std::isfinite(th_t)) {
            // ---- second-order correction
            double a_soc = alpha; double th_prev = th_t;
            for (int r = 0; r < m; ++r) csoc[r] = c[r];
            std::vector<double> gs = gt, ss = st;
            for (int k = 0; k < o.max_soc; ++k) {
              for (int r = 0; r < m; ++r) csoc[r] = a_soc * csoc[r] + (gs[r] - ss[r]);
              solve_dir(Ds, rx, rs, csoc.data(), dx2.data(), ds2.data(), dy2.data());
              a_soc = ftb_primal(dx2.data(), ds2.data());
              for (int i = 0; i < n; ++i) xt[i] = x[i] + a_soc * dx2[i];
              for (int r = 0; r < m; ++r) ss[r] = s[r] + a_soc * ds2[r];
              double f_s = feval(xt.data(), gs.data());
              double th_s = 0; for (int r = 0; r < m; ++r) th_s += std::fabs(gs[r] - ss[r]);
              double ph_s = barrier(xt.data(), ss.data(), f_s);
              ++ls;
              if (acceptable(alpha, th_s, ph_s)) {
                accepted = true; f_t = f_s; gt = gs; st = ss; alpha = a_soc;
                use_dx = dx2.data(); use_ds = ds2.data(); use_dy = dy2.data(); ++n_soc_acc;
                break;
              }
              if (!(th_s <= o.kappa_soc * th_prev)) break;
              th_prev = th_s;
            }
            if (accepted) break;
          }
          first = false;
          alpha *= o.alpha_red;
        }
      }
      if (status == STEP_TOO_SMALL) break;
      if (!accepted) { status = RESTORATION_NEEDED; break; }
      // -------- filter augmentation (uses the reference point and the original direction)
      if (!tiny && !(is_ftype(alpha_test) && [&] {
            double ph_t = barrier(xt.data(), st.data(), f_t); return armijo(alpha_test, ph_t); }())) {
        filter.emplace_back((1 - o.gamma_theta) * theta, phi - o.gamma_phi * theta);
      }
      // -------- accept
      if (use_dx != dx.data()) { dual_dirs(use_dx, use_ds); a_du = ftb_dual(); }
      if (log) log->push_back({mu, f / df, pr, du, dw, alpha, a_du, ls});
      x = xt; s = st; g = gt; f = f_t;
      for (int r = 0; r < m; ++r) y[r] += alpha * use_dy[r];
      for (int i = 0; i < n; ++i) { zL[i] += a_du * dzL[i]; zU[i] += a_du * dzU[i]; }
      for (int r = 0; r < m; ++r) { vL[r] += a_du * dvL[r]; vU[r] += a_du * dvU[r]; }
      auto reset = [&](double& z, double sl) { z = std::max(std::min(z, o.kappa_sigma * mu / sl), mu / (o.kappa_sigma * sl)); };
      for (int i = 0; i < n; ++i) { if (xL[i] > -INF) reset(zL[i], x[i] - xL[i]); if (xU[i] < INF) reset(zU[i], xU[i] - x[i]); }
      for (int r = 0; r < m; ++r) { if (dL[r] > -INF) reset(vL[r], s[r] - dL[r]); if (dU[r] < INF) reset(vU[r], dU[r] - s[r]); }
      deriv(x.data());
      ++iter;
    }
    // ---- finalize: honour original bounds, unscale
    for (int i = 0; i < n; ++i) x_out[i] = std::min(std::max(x[i], lbx[i]), ubx[i]);
    std::vector<double> gu(m);
    double fu = P.eval_fg(x_out, gu.data());   // unscaled f, g at the returned point
    if (g_out) std::memcpy(g_out, gu.data(), sizeof(double) * m);
    if (lamx_out) for (int i = 0; i < n; ++i) lamx_out[i] = (zU[i] - zL[i]) / df;
    if (lamg_out) for (int r = 0; r < m; ++r) lamg_out[r] = y[r] * dc[r] / df;
    res.status = status; res.iters = iter; res.f = fu; res.n_fact = n_fact; res.n_soc_acc = n_soc_acc;
    return res;
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------
// C interface for ctypes (tests / bench only)
// ------------------------------------------------------------------------------------------
extern "C" {

struct oracle_spec { double T; int32_t N; int32_t n_obs; double w1, w2, vfov, hfov; };

static Spec to_spec(const oracle_spec* s) { return Spec{s->T, s->N, s->n_obs, s->w1, s->w2, s->vfov, s->hfov}; }

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
// stats (optional) [B][4] : n_fact, n_soc_accepted, 0, 0
int oracle_solve_traj(const oracle_spec* spec, int B, const double* p, const double* x0, const double* tgt,
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
  Spec sp = to_spec(spec); const int nw = sp.nw(), ng = sp.ng();
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      int b = next.fetch_add(1); if (b >= B) break;
      const double* ob = obs + (obs_per_instance ? (size_t)b * sp.n_obs * 3 : 0);
      Instance I(sp, p + (size_t)b * NPAR, ob);
      if (tgt) I.tgt = tgt + (size_t)b * 2 * sp.N;
      Options o; o.scaling = scaling; if (max_iter > 0) o.max_iter = max_iter; if (tol > 0) o.tol = tol;
      Ipm ipm(I, o);
      std::vector<double> xo(nw);
      Result r = ipm.solve(x0 + (size_t)b * nw, lbx, ubx, lbg, ubg, xo.data(),
                           g ? g + (size_t)b * ng : nullptr, lam_x ? lam_x + (size_t)b * nw : nullptr,
                           lam_g ? lam_g + (size_t)b * ng : nullptr);
      std::memcpy(x + (size_t)b * nw, xo.data(), sizeof(double) * nw);
      if (f) f[b] = r.f;
      if (status) status[b] = r.status;
      if (iters) iters[b] = r.iters;
      if (stats) { stats[4 * b] = r.n_fact; stats[4 * b + 1] = r.n_soc_acc; stats[4 * b + 2] = 0; stats[4 * b + 3] = 0; }
    }
  };
  if (nthreads == 1) work();
  else { std::vector<std::thread> th; for (int t = 0; t < nthreads; ++t) th.emplace_back(work); for (auto& t : th) t.join(); }
  return 0;
}

// single solve with per-iteration log [max_log][8]: mu, f, inf_pr, inf_du, dw, alpha_pr, alpha_du, ls
int oracle_solve_log(const oracle_spec* spec, const double* p, const double* x0,
                     const double* lbx, const double* ubx, const double* lbg, const double* ubg,
                     const double* obs, int scaling, double* x, double* f, int32_t* status, int32_t* iters,
                     double* logbuf, int max_log) {
  Spec sp = to_spec(spec); Instance I(sp, p, obs);
  Options o; o.scaling = scaling; Ipm ipm(I, o);
  std::vector<IterLog> lg; ipm.log = &lg;
  Result r = ipm.solve(x0, lbx, ubx, lbg, ubg, x, nullptr, nullptr, nullptr);
  *f = r.f; *status = r.status; *iters = r.iters;
  for (int i = 0; i < (int)lg.size() && i < max_log; ++i) {
    double* L = logbuf + 8 * i;
    L[0] = lg[i].mu; L[1] = lg[i].f; L[2] = lg[i].inf_pr; L[3] = lg[i].inf_du; L[4] = lg[i].dw;
    L[5] = lg[i].alpha_pr; L[6] = lg[i].alpha_du; L[7] = lg[i].ls;
  }
  return (int)lg.size();
}

}  // extern "C"
