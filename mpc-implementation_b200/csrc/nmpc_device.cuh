// nmpc_device.cuh -- device-side building blocks of the batched NMPC solver (sm_100a, FP64).
//
// One warp owns one NLP instance; lane k owns horizon stage k (N + 1 <= 32).
//   * rollout (NMPC_TT.py:160-167) = warp prefix sums: the five angle states are pure integrators
//     of their rate controls, the positions are prefix sums of T*v*[cos(psi)cos(th), sin(psi)cos(th), sin(th)].
//   * stage cost (NMPC_TT.py:193-221) in its compact form  w1*dist + w2*((P/a)^2 + (Q/b)^2 - 1)
//     with hand-derived gradient and Hessian (chain rule through the intermediates (Dx, Dy, a, b, X7)).
//   * Newton step of the reduced (single-shooting) primal-dual system = stage-structured LQ problem,
//     solved by a Riccati recursion whose 8x8 / 6x8 / 6x6 blocks are spread over lanes 0..8
//     (lane j = column j of [P | p]), with warp shuffles for the cross-column terms.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace nmpc {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NX = 8, NU = 6, NPAR = 11;

// ---- per-row (constraint) arrays kept in shared memory: index (arr * R + r) * S + stage -------
enum RowArr { A_S = 0, A_Y, A_VL, A_VU, A_G, A_GT, A_DS, A_DS2, A_CSOC, A_DC, A_DL, A_DU, A_NROW };

// ---- per-stage LQ data (entry-major: e * S + stage) ------------------------------------------
enum LqEnt {
  LQ_Q = 0,        // 36: packed lower triangle of the 8x8 state Hessian block
  LQ_SV = 36,      // 2 : d2/dv dtheta, d2/dv dpsi (only non-zeros of the control-state block)
  LQ_RD = 38,      // 6 : diagonal control block (Sigma_x, without delta_w)
  LQ_QV = 44,      // 8 : state gradient q
  LQ_RV = 52,      // 6 : control gradient r
  LQ_DD = 58,      // 3 : direction d = (cps*cth, sps*cth, sth)
  LQ_EE = 61,      // 5 : T*E03, T*E13, T*E23, T*E04, T*E14
  LQ_NN = 66,      // 3 : sum_j dc_j^2 n_j n_j^T  (xx, xy, yy)   -- delta_w part of G^T D_s G
  LQ_DG = 69,      // 5 : dc_r^2 for the five box rows            -- delta_w part of G^T D_s G
  LQ_Q2 = 74,      // 8 : SOC state gradient q'
  LQ_QD = 82,      // 8 : G^T c  -- delta_w part of the state gradient q
  LQ_N = 90
};
constexpr int RIC_K = 0;      // 54: K (6x8) and kappa as a 6x9 row-major block
constexpr int RIC_L = 54;     // 21: Cholesky factor of Lambda, packed lower
constexpr int RIC_K2 = 75;    // 6 : kappa of the SOC right-hand side
constexpr int RIC_N = 81;
constexpr int FILT_CAP = 24;

__host__ __device__ inline int tri(int r, int c) { return r >= c ? r * (r + 1) / 2 + c : c * (c + 1) / 2 + r; }
// state index of the five linear ("box") rows [z, theta, X5, X6, X7]   NMPC_TT.py:236-240
__device__ __forceinline__ int box_state(int r) { return r == 0 ? 2 : (r == 1 ? 3 : r + 3); }

// ---- warp collectives ------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double scan_excl(double v, int lane) {   // exclusive prefix sum over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { double t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
  double e = __shfl_up_sync(FULL, v, 1);
  return lane == 0 ? 0.0 : e;
}
__device__ __forceinline__ double rscan_incl(double v, int lane) {  // inclusive suffix sum over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { double t = __shfl_down_sync(FULL, v, o); if (lane + o < 32) v += t; }
  return v;
}
__device__ __forceinline__ double shfl_next(double v, int lane) {   // value of lane+1 (0 for the last lane)
  double t = __shfl_down_sync(FULL, v, 1);
  return lane == 31 ? 0.0 : t;
}

// ---- problem constants passed by value to the kernels ----------------------------------------
struct Prob {
  double T, w1, w2, hv, hh;   // hv = VFOV/2, hh = HFOV/2
  int N, n_obs, R, S;         // R = 5 + n_obs rows per stage, S = N + 1 stages
};

// ---- stage state of one lane -----------------------------------------------------------------
struct Stage {
  double X[NX];
  double sth, cth, sps, cps;
};

// Rollout by prefix sums.  u is this lane's control (zero for lanes >= N).  NMPC_TT.py:139-148,160-167
__device__ __forceinline__ void rollout(const Prob& pr, const double* X0, const double* u, int lane, Stage& st) {
  const bool has_u = lane < pr.N;
#pragma unroll
  for (int c = 0; c < 5; ++c) st.X[3 + c] = X0[3 + c] + scan_excl(has_u ? pr.T * u[c + 1] : 0.0, lane);
  sincos(st.X[3], &st.sth, &st.cth);
  sincos(st.X[4], &st.sps, &st.cps);
  const double tv = has_u ? pr.T * u[0] : 0.0;
  st.X[0] = X0[0] + scan_excl(tv * st.cps * st.cth, lane);
  st.X[1] = X0[1] + scan_excl(tv * st.sps * st.cth, lane);
  st.X[2] = X0[2] + scan_excl(tv * st.sth, lane);
}

// FOV geometry shared by value / derivative evaluation
struct Fov {
  double t6p, t6m, t5p, t5m, s7, c7;
};
__device__ __forceinline__ void fov_trig(const Prob& pr, const double* X, Fov& f) {
  f.t6p = tan(X[6] + pr.hv); f.t6m = tan(X[6] - pr.hv);
  f.t5p = tan(X[5] + pr.hh); f.t5m = tan(X[5] - pr.hh);
  sincos(X[7], &f.s7, &f.c7);
}

// stage cost value (compact form of NMPC_TT.py:209-220)
__device__ __forceinline__ double stage_cost(const Prob& pr, const double* X, double xt, double yt) {
  Fov f; fov_trig(pr, X, f);
  const double z = X[2];
  const double a = z * 0.5 * (f.t6p - f.t6m), b = z * 0.5 * (f.t5p - f.t5m);
  const double Dx = xt - X[0] - z * 0.5 * (f.t6p + f.t6m);
  const double Dy = yt - X[1] - z * 0.5 * (f.t5p + f.t5m);
  const double U = (f.c7 * Dx + f.s7 * Dy) / a, V = (f.s7 * Dx - f.c7 * Dy) / b;
  const double ex = X[0] - xt, ey = X[1] - yt;
  return pr.w1 * sqrt(ex * ex + ey * ey) + pr.w2 * (U * U + V * V - 1.0);
}

// stage cost with gradient gl[6] and packed Hessian Hl[21] over (x, y, z, X5, X6, X7)
__device__ __forceinline__ double stage_cost_d2(const Prob& pr, const double* X, double xt, double yt, double* gl, double* Hl) {
  Fov f; fov_trig(pr, X, f);
  const double z = X[2];
  const double d6p = 1.0 + f.t6p * f.t6p, d6m = 1.0 + f.t6m * f.t6m, d5p = 1.0 + f.t5p * f.t5p, d5m = 1.0 + f.t5m * f.t5m;
  const double e6p = 2.0 * f.t6p * d6p, e6m = 2.0 * f.t6m * d6m, e5p = 2.0 * f.t5p * d5p, e5m = 2.0 * f.t5m * d5m;
  const double al = 0.5 * (f.t6p - f.t6m), al1 = 0.5 * (d6p - d6m), al2 = 0.5 * (e6p - e6m);
  const double ga = 0.5 * (f.t6p + f.t6m), ga1 = 0.5 * (d6p + d6m), ga2 = 0.5 * (e6p + e6m);
  const double be = 0.5 * (f.t5p - f.t5m), be1 = 0.5 * (d5p - d5m), be2 = 0.5 * (e5p - e5m);
  const double ep = 0.5 * (f.t5p + f.t5m), ep1 = 0.5 * (d5p + d5m), ep2 = 0.5 * (e5p + e5m);
  const double a = z * al, b = z * be;
  const double Dx = xt - X[0] - z * ga, Dy = yt - X[1] - z * ep;
  const double s7 = f.s7, c7 = f.c7;
  const double Ph = c7 * Dx + s7 * Dy, Qh = s7 * Dx - c7 * Dy;
  const double ia = 1.0 / a, ib = 1.0 / b;
  const double U = Ph * ia, V = Qh * ib;
  // intermediates m = (Dx, Dy, a, b, X7)
  const double Um[5] = {c7 * ia, s7 * ia, -U * ia, 0.0, -Qh * ia};
  const double Vm[5] = {s7 * ib, -c7 * ib, 0.0, -V * ib, Ph * ib};
  double gm[5], Hm[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i) gm[i] = 2.0 * (U * Um[i] + V * Vm[i]);
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) Hm[i][j] = 2.0 * (Um[i] * Um[j] + Vm[i] * Vm[j]);
  const double U2 = 2.0 * U, V2 = 2.0 * V, ia2 = ia * ia, ib2 = ib * ib;
  auto addsym = [&](int i, int j, double v) { Hm[i][j] += v; if (i != j) Hm[j][i] += v; };
  addsym(0, 2, U2 * (-c7 * ia2)); addsym(1, 2, U2 * (-s7 * ia2)); addsym(2, 2, U2 * (2.0 * U * ia2));
  addsym(0, 4, U2 * (-s7 * ia)); addsym(1, 4, U2 * (c7 * ia)); addsym(2, 4, U2 * (Qh * ia2)); addsym(4, 4, U2 * (-U));
  addsym(0, 3, V2 * (-s7 * ib2)); addsym(1, 3, V2 * (c7 * ib2)); addsym(3, 3, V2 * (2.0 * V * ib2));
  addsym(0, 4, V2 * (c7 * ib)); addsym(1, 4, V2 * (s7 * ib)); addsym(3, 4, V2 * (-Ph * ib2)); addsym(4, 4, V2 * (-V));
  // Jacobian of m w.r.t. v = (x, y, z, X5, X6, X7)
  const double Jm[5][6] = {{-1.0, 0.0, -ga, 0.0, -z * ga1, 0.0},
                           {0.0, -1.0, -ep, -z * ep1, 0.0, 0.0},
                           {0.0, 0.0, al, 0.0, z * al1, 0.0},
                           {0.0, 0.0, be, z * be1, 0.0, 0.0},
                           {0.0, 0.0, 0.0, 0.0, 0.0, 1.0}};
  double HJ[5][6];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 5; ++j) acc += Hm[i][j] * Jm[j][v];
      HJ[i][v] = acc;
    }
  const double w2 = pr.w2;
#pragma unroll
  for (int v = 0; v < 6; ++v) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) acc += gm[i] * Jm[i][v];
    gl[v] = w2 * acc;
  }
#pragma unroll
  for (int u = 0; u < 6; ++u)
#pragma unroll
    for (int v = 0; v <= u; ++v) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < 5; ++i) acc += Jm[i][u] * HJ[i][v];
      Hl[tri(u, v)] = w2 * acc;
    }
  // curvature of the intermediates themselves
  Hl[tri(4, 2)] += w2 * (gm[0] * (-ga1) + gm[2] * al1);          // (z, X6)
  Hl[tri(4, 4)] += w2 * (gm[0] * (-z * ga2) + gm[2] * (z * al2)); // (X6, X6)
  Hl[tri(3, 2)] += w2 * (gm[1] * (-ep1) + gm[3] * be1);          // (z, X5)
  Hl[tri(3, 3)] += w2 * (gm[1] * (-z * ep2) + gm[3] * (z * be2)); // (X5, X5)
  // distance term
  const double ex = X[0] - xt, ey = X[1] - yt;
  const double D = sqrt(ex * ex + ey * ey), iD = 1.0 / D;
  const double nx = ex * iD, ny = ey * iD, w1 = pr.w1;
  gl[0] += w1 * nx; gl[1] += w1 * ny;
  Hl[tri(0, 0)] += w1 * iD * (1.0 - nx * nx);
  Hl[tri(1, 0)] += w1 * iD * (-nx * ny);
  Hl[tri(1, 1)] += w1 * iD * (1.0 - ny * ny);
  return w1 * D + w2 * (U * U + V * V - 1.0);
}

// compact cost variable (x,y,z,X5,X6,X7) -> state index
__device__ __forceinline__ int cost_state(int v) { return v < 3 ? v : v + 2; }

// 6x6 Cholesky on a packed lower triangle (in place).  Returns false when a pivot is not positive.
__device__ __forceinline__ bool chol6(double* L, double* inv_diag) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = L[tri(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[tri(j, k)] * L[tri(j, k)];
    ok = ok && (d > 0.0);
    const double sd = sqrt(d), id = 1.0 / sd;
    L[tri(j, j)] = sd; inv_diag[j] = id;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double s = L[tri(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[tri(i, k)] * L[tri(j, k)];
      L[tri(i, j)] = s * id;
    }
  }
  return ok;
}
__device__ __forceinline__ void chol6_solve(const double* L, const double* inv_diag, double* b) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[tri(i, k)] * b[k];
    b[i] = s * inv_diag[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = b[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= L[tri(k, i)] * b[k];
    b[i] = s * inv_diag[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Riccati factorisation + solve of the stage-structured Newton system (all lanes participate).
//   lane j < 8 : column j of P and of Psi / K;   lane 8 : the vector column (p, psi -> kappa).
// Returns false as soon as a pivot block Lambda_k is not positive definite (wrong inertia).
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ bool riccati_factor(const Prob& pr, const double* __restrict__ lq, double* __restrict__ ric,
                                            double dw, int lane) {
  const int S = pr.S, N = pr.N;
  const int j = lane < 8 ? lane : 8;
  const double T = pr.T;
  auto Qent = [&](int k, int i) -> double {   // Q_k[i][j] including the delta_w part; q_k[i] for the vector lane
    if (j == 8) return lq[(LQ_QV + i) * S + k] + dw * lq[(LQ_QD + i) * S + k];
    double v = lq[(LQ_Q + tri(i, j)) * S + k];
    if (dw != 0.0) {
      if (i < 2 && j < 2) v += dw * lq[(LQ_NN + i + j) * S + k];
      else if (i == j && i != 4) { const int r = (i == 2) ? 0 : (i == 3 ? 1 : i - 3); v += dw * lq[(LQ_DG + r) * S + k]; }
    }
    return v;
  };
  double P[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) P[i] = Qent(N, i);
  for (int k = N - 1; k >= 0; --k) {
    const double d0 = lq[(LQ_DD + 0) * S + k], d1 = lq[(LQ_DD + 1) * S + k], d2 = lq[(LQ_DD + 2) * S + k];
    const double e03 = lq[(LQ_EE + 0) * S + k], e13 = lq[(LQ_EE + 1) * S + k], e23 = lq[(LQ_EE + 2) * S + k];
    const double e04 = lq[(LQ_EE + 3) * S + k], e14 = lq[(LQ_EE + 4) * S + k];
    // M = P+ A  (column j)
    double M[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double c0 = __shfl_sync(FULL, P[i], 0), c1 = __shfl_sync(FULL, P[i], 1), c2 = __shfl_sync(FULL, P[i], 2);
      double add = 0.0;
      if (j == 3) add = e03 * c0 + e13 * c1 + e23 * c2;
      else if (j == 4) add = e04 * c0 + e14 * c1;
      M[i] = P[i] + add;
    }
    // G = B^T P+ (column j),  Psi = S + B^T M (column j; the vector lane adds r)
    double G[6], Psi[6];
    G[0] = T * (d0 * P[0] + d1 * P[1] + d2 * P[2]);
    Psi[0] = T * (d0 * M[0] + d1 * M[1] + d2 * M[2]);
#pragma unroll
    for (int r = 1; r < 6; ++r) { G[r] = T * P[r + 2]; Psi[r] = T * M[r + 2]; }
    if (j == 3) Psi[0] += lq[(LQ_SV + 0) * S + k];
    if (j == 4) Psi[0] += lq[(LQ_SV + 1) * S + k];
    if (j == 8) {
#pragma unroll
      for (int r = 0; r < 6; ++r) Psi[r] += lq[(LQ_RV + r) * S + k];
    }
    // A^T M (column j)
    double AtM[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) AtM[i] = M[i];
    AtM[3] += e03 * M[0] + e13 * M[1] + e23 * M[2];
    AtM[4] += e04 * M[0] + e14 * M[1];
    // Lambda = R + B^T P+ B  gathered to every lane (packed lower)
    double L[21], idg[6];
#pragma unroll
    for (int c = 1; c < 6; ++c)
#pragma unroll
      for (int r = c; r < 6; ++r) L[tri(r, c)] = T * __shfl_sync(FULL, G[r], c + 2);
#pragma unroll
    for (int r = 1; r < 6; ++r) L[tri(r, 0)] = T * __shfl_sync(FULL, G[0], r + 2);
    L[0] = T * (d0 * __shfl_sync(FULL, G[0], 0) + d1 * __shfl_sync(FULL, G[0], 1) + d2 * __shfl_sync(FULL, G[0], 2));
#pragma unroll
    for (int r = 0; r < 6; ++r) L[tri(r, r)] += lq[(LQ_RD + r) * S + k] + dw;
    if (!chol6(L, idg)) return false;
    // K column = Lambda^-1 Psi column
    double Kc[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) Kc[r] = Psi[r];
    chol6_solve(L, idg, Kc);
    // P = Q + A^T P+ A - Psi^T K
    double Pn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double acc = Qent(k, i) + AtM[i];
#pragma unroll
      for (int r = 0; r < 6; ++r) acc -= __shfl_sync(FULL, Psi[r], i) * Kc[r];
      Pn[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) P[i] = Pn[i];
    double* rk = ric + k * RIC_N;
    if (lane < 9) {
#pragma unroll
      for (int r = 0; r < 6; ++r) rk[RIC_K + r * 9 + j] = Kc[r];
    }
    if (lane == 9) {
#pragma unroll
      for (int e = 0; e < 21; ++e) rk[RIC_L + e] = L[e];
    }
  }
  __syncwarp();
  return true;
}

// Forward sweep: du_k = -K_k dx_k - kappa_k, dx_{k+1} = A_k dx_k + B_k du_k (computed redundantly by all lanes;
// lane k keeps its own stage's dx, du).  kap_off selects kappa (RIC_K column 8) or the SOC kappa (RIC_K2).
__device__ __noinline__ void riccati_forward(const Prob& pr, const double* __restrict__ lq, const double* __restrict__ ric,
                                             bool soc, int lane, double* mydx, double* mydu) {
  const int S = pr.S, N = pr.N; const double T = pr.T;
  double dx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dx[i] = 0.0; mydx[i] = 0.0; }
#pragma unroll
  for (int r = 0; r < 6; ++r) mydu[r] = 0.0;
  for (int k = 0; k < N; ++k) {
    const double* rk = ric + k * RIC_N;
    double du[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double acc = soc ? rk[RIC_K2 + r] : rk[RIC_K + r * 9 + 8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += rk[RIC_K + r * 9 + i] * dx[i];
      du[r] = -acc;
    }
    if (lane == k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mydx[i] = dx[i];
#pragma unroll
      for (int r = 0; r < 6; ++r) mydu[r] = du[r];
    }
    const double d0 = lq[(LQ_DD + 0) * S + k], d1 = lq[(LQ_DD + 1) * S + k], d2 = lq[(LQ_DD + 2) * S + k];
    const double e03 = lq[(LQ_EE + 0) * S + k], e13 = lq[(LQ_EE + 1) * S + k], e23 = lq[(LQ_EE + 2) * S + k];
    const double e04 = lq[(LQ_EE + 3) * S + k], e14 = lq[(LQ_EE + 4) * S + k];
    const double tv = T * du[0];
    dx[0] += e03 * dx[3] + e04 * dx[4] + tv * d0;
    dx[1] += e13 * dx[3] + e14 * dx[4] + tv * d1;
    dx[2] += e23 * dx[3] + tv * d2;
#pragma unroll
    for (int r = 1; r < 6; ++r) dx[r + 2] += T * du[r];
  }
  if (lane == N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) mydx[i] = dx[i];
  }
}

// Backward sweep for a new state-gradient q' (LQ_Q2) with the stored factors (SOC right-hand sides).
__device__ __noinline__ void riccati_resolve(const Prob& pr, const double* __restrict__ lq, double* __restrict__ ric, int lane) {
  const int S = pr.S, N = pr.N; const double T = pr.T;
  double p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = lq[(LQ_Q2 + i) * S + N];
  for (int k = N - 1; k >= 0; --k) {
    double* rk = ric + k * RIC_N;
    const double d0 = lq[(LQ_DD + 0) * S + k], d1 = lq[(LQ_DD + 1) * S + k], d2 = lq[(LQ_DD + 2) * S + k];
    const double e03 = lq[(LQ_EE + 0) * S + k], e13 = lq[(LQ_EE + 1) * S + k], e23 = lq[(LQ_EE + 2) * S + k];
    const double e04 = lq[(LQ_EE + 3) * S + k], e14 = lq[(LQ_EE + 4) * S + k];
    double psi[6], kap[6];
    psi[0] = lq[(LQ_RV + 0) * S + k] + T * (d0 * p[0] + d1 * p[1] + d2 * p[2]);
#pragma unroll
    for (int r = 1; r < 6; ++r) psi[r] = lq[(LQ_RV + r) * S + k] + T * p[r + 2];
    double L[21], idg[6];
#pragma unroll
    for (int e = 0; e < 21; ++e) L[e] = rk[RIC_L + e];
#pragma unroll
    for (int r = 0; r < 6; ++r) { idg[r] = 1.0 / L[tri(r, r)]; kap[r] = psi[r]; }
    chol6_solve(L, idg, kap);
    double pn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pn[i] = lq[(LQ_Q2 + i) * S + k] + p[i];
    pn[3] += e03 * p[0] + e13 * p[1] + e23 * p[2];
    pn[4] += e04 * p[0] + e14 * p[1];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) acc += rk[RIC_K + r * 9 + i] * psi[r];
      pn[i] -= acc;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pn[i];
    __syncwarp();
    if (lane < 6) rk[RIC_K2 + lane] = kap[lane];
  }
  __syncwarp();
}

}  // namespace nmpc
