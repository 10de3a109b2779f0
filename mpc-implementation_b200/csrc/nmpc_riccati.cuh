// nmpc_riccati.cuh -- warp-parallel Riccati factorisation / solves of the stage-structured Newton system.
//
// Per stage k (backward) the 15 x 15 symmetric "stage KKT matrix" in the ordering (u (6), x (8), rhs (1)) is
//     H = Z^T P+ Z + [R S; S^T Q],   h = Z^T p+ + [r; q],          Z = [B_k | A_k]  (8 x 14),
// and eliminating the six control pivots by a right-looking Cholesky leaves P_k (8 x 8) and p_k in the trailing
// block, L (6 x 6), L^-1 Psi (6 x 8) and L^-1 psi in the pivot columns.  All 32 lanes work: every lane OWNS up
// to four entries of the packed lower triangle (registers), operands are exchanged through a 3.4 KB shared-memory
// staging area, and every loop is rolled so the whole stage body stays resident in the L0 instruction cache
// (the first version was 1900 unrolled instructions with 9 of 32 lanes active and stalled on instruction fetch).
// "Inertia correct" (IPOPT)  <=>  all six pivots of every stage positive.
#pragma once
#include "nmpc_device.cuh"

namespace nmpc {

// staging area layout (doubles)
constexpr int SG_PP = 0;      // [9][8]  rows 0..7 = P+ (symmetric), row 8 = p+
constexpr int SG_Z = 72;      // [8][16] Z[s][j], j < 14
constexpr int SG_Y = 200;     // [15][8] Y_j = P+ z_j (j < 14), Y_14 = p+
constexpr int SG_COL = 320;   // [16]    published pivot column
constexpr int SG_FAC = 336;   // [6][16] factor columns l[a], a = 0..14, and 1/l[j][j] at [15]
constexpr int SG_N = 432;

// state-Hessian entry Q_k[i][j] (+ delta_w part) / gradient entry, used for the terminal stage only
__device__ __forceinline__ double q_entry(const double* lq, int S, int k, int i, int j, double dw) {
  const int qi = q_index(i, j);
  double v = qi >= 0 ? lq[(LQ_Q + qi) * S + k] : 0.0;
  if (i < 2 && j < 2) v += dw * lq[(LQ_NN + i + j) * S + k];
  else if (i == j && i != 4) v += dw * lq[(LQ_DG + ((i == 2) ? 0 : (i == 3 ? 1 : i - 3))) * S + k];
  return v;
}

__device__ __noinline__ bool riccati_factor(const Prob& pr, const double* lq, double* ric, double* stg, double mu, double dw, int lane) {
  const int S = pr.S, N = pr.N;
  const double T = pr.T;
  double* PP = stg + SG_PP; double* ZS = stg + SG_Z; double* YS = stg + SG_Y; double* COL = stg + SG_COL; double* FAC = stg + SG_FAC;
  // ---- ownership map: entry e = lane + 32 m of the packed lower triangle of the 15 x 15 matrix, and the
  //      LQ entries that are added to it:  ent += lq[o0] + mu * lq[o1] + dw * (lq[o2] + c3)
  int ia[4], ib[4], o0[4], o1[4], o2[4]; double c3[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int e = lane + 32 * m;
    int a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while (a * (a + 1) / 2 > e) --a;
    while ((a + 1) * (a + 2) / 2 <= e) ++a;
    const int b = e - a * (a + 1) / 2;
    const bool valid = e < 119;                       // e == 119 is (14,14): not needed
    ia[m] = valid ? a : -1; ib[m] = valid ? b : 99;
    int p0 = LQ_ZERO, p1 = LQ_ZERO, p2 = LQ_ZERO; double cc = 0.0;
    if (valid) {
      if (a < 6) { if (a == b) { p0 = LQ_RD + a; cc = 1.0; } }
      else if (a < 14) {
        if (b < 6) { if (b == 0 && a == 9) p0 = LQ_SV + 0; else if (b == 0 && a == 10) p0 = LQ_SV + 1; }
        else {
          const int i = a - 6, j = b - 6, qi = q_index(i, j);
          if (qi >= 0) p0 = LQ_Q + qi;
          if (i < 2 && j < 2) p2 = LQ_NN + i + j;
          else if (i == j && i != 4) p2 = LQ_DG + ((i == 2) ? 0 : (i == 3 ? 1 : i - 3));
        }
      } else {
        if (b < 6) p1 = LQ_RB + b;
        else { p0 = LQ_QA + (b - 6); p1 = LQ_QB + (b - 6); p2 = LQ_QD + (b - 6); }
      }
    }
    o0[m] = p0 * S; o1[m] = p1 * S; o2[m] = p2 * S; c3[m] = cc;
  }
  // ---- static part of Z = [B | A]:  B = T [d e_v ; I_5 on the angle rows],  A = I + E
  for (int o = lane; o < 128; o += 32) {
    const int s = o >> 4, j = o & 15;
    double v = 0.0;
    if (j >= 1 && j < 6 && s == j + 2) v = T;
    if (j >= 6 && j < 14 && s == j - 6) v = 1.0;
    ZS[o] = v;
  }
  // ---- terminal stage: P_N = Q_N, p_N = q_N
  for (int o = lane; o < 72; o += 32) {
    const int i = o >> 3, j = o & 7;
    PP[o] = i < 8 ? q_entry(lq, S, N, i, j, dw)
                  : lq[(LQ_QA + j) * S + N] + mu * lq[(LQ_QB + j) * S + N] + dw * lq[(LQ_QD + j) * S + N];
  }
  __syncwarp();
  for (int k = N - 1; k >= 0; --k) {
    // stage-dependent entries of Z
    if (lane < 8) {
      const int t = lane;
      const int dst = t < 3 ? t * 16 : (t < 6 ? (t - 3) * 16 + 9 : (t - 6) * 16 + 10);
      ZS[dst] = t < 3 ? T * lq[(LQ_DD + t) * S + k] : lq[(LQ_EE + (t - 3)) * S + k];
    }
    __syncwarp();
    // Y_j = P+ z_j  (j < 14),  Y_14 = p+
#pragma unroll 1
    for (int m = 0; m < 4; ++m) {
      const int o = lane + 32 * m, j = o >> 3, i = o & 7;
      if (j < 14) {
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < 8; ++s) acc += ZS[s * 16 + j] * PP[i * 8 + s];
        YS[o] = acc;
      } else if (j == 14) YS[o] = PP[64 + i];
    }
    __syncwarp();
    // owned entries of [H | h]
    double ent[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double acc = 0.0;
      if (ia[m] >= 0) {
        const int zc = ia[m] == 14 ? ib[m] : ia[m], yr = ia[m] == 14 ? 14 : ib[m];
#pragma unroll
        for (int s = 0; s < 8; ++s) acc += ZS[s * 16 + zc] * YS[yr * 8 + s];
        acc += lq[o0[m] + k] + mu * lq[o1[m] + k] + dw * (lq[o2[m] + k] + c3[m]);
      }
      ent[m] = acc;
    }
    // partial Cholesky: eliminate the six control pivots
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {
#pragma unroll
      for (int m = 0; m < 4; ++m) if (ib[m] == j) COL[ia[m]] = ent[m];
      __syncwarp();
      const double d = COL[j];
      if (!(d > 0.0)) return false;                 // uniform: every lane reads the same pivot
      const double id = rsqrt(d);
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (ib[m] > j && ib[m] < 99) ent[m] -= (COL[ia[m]] * id) * (COL[ib[m]] * id);
      if (lane >= j && lane < 15) FAC[j * 16 + lane] = COL[lane] * id;
      if (lane == 15) FAC[j * 16 + 15] = id;
      __syncwarp();
    }
    // trailing block -> P_k, p_k for the next stage
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (ib[m] >= 6 && ib[m] < 99) {
        const int a = ia[m] - 6, b = ib[m] - 6;
        if (ia[m] == 14) PP[64 + b] = ent[m];
        else { PP[a * 8 + b] = ent[m]; PP[b * 8 + a] = ent[m]; }
      }
    }
    // K = L^-T [L^-1 Psi | L^-1 psi]: lane c < 9 back-substitutes column c
    double* rk = ric + k * RIC_N;
    if (lane < 9) {
      double x[6];
#pragma unroll
      for (int i = 5; i >= 0; --i) {
        double s = FAC[i * 16 + 6 + lane];
#pragma unroll
        for (int m = i + 1; m < 6; ++m) s -= FAC[i * 16 + m] * x[m];
        x[i] = s * FAC[i * 16 + 15];
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) rk[RIC_K + r * 9 + lane] = x[r];
    } else if (lane < 30) {            // L (packed lower) for the SOC re-solves: entry e = tri(r, c) = FAC[c][r]
      const int e = lane - 9;
      const int r = (e >= 1) + (e >= 3) + (e >= 6) + (e >= 10) + (e >= 15), c = e - r * (r + 1) / 2;
      rk[RIC_L + e] = FAC[c * 16 + r];
    }
    if (lane < 6) rk[RIC_L + 21 + lane] = FAC[lane * 16 + 15];
    __syncwarp();
  }
  return true;
}

// Forward sweep: du_k = -K_k dx_k - kappa_k, dx_{k+1} = A_k dx_k + B_k du_k.  Lane r < 6 computes row r of du
// from its row of [K | kappa] (prefetched one stage ahead from the L2-resident scratch); dx is replicated.
// Lane k keeps its own stage's dx, du and writes them to dxo[8][S], duo[6][S].
__device__ __noinline__ void riccati_forward(const Prob& pr, const double* lq, const double* ric,
                                             bool soc, int lane, double* dxo, double* duo) {
  const int S = pr.S, N = pr.N; const double T = pr.T;
  const int r = lane < 6 ? lane : 5;
  double dx[8], mydx[8], mydu[6];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dx[i] = 0.0; mydx[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < 6; ++i) mydu[i] = 0.0;
  double row[9];
  auto load_row = [&](int k) {
    const double* rk = ric + k * RIC_N;
#pragma unroll
    for (int i = 0; i < 8; ++i) row[i] = __ldcg(rk + RIC_K + r * 9 + i);   // written by this kernel: never the .nc path
    row[8] = soc ? __ldcg(rk + RIC_K2 + r) : __ldcg(rk + RIC_K + r * 9 + 8);
  };
  load_row(0);
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    double acc = row[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += row[i] * dx[i];
    if (k + 1 < N) load_row(k + 1);
    double du[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) du[i] = -__shfl_sync(FULL, acc, i);
    if (lane == k) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mydx[i] = dx[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) mydu[i] = du[i];
    }
    const double d0 = lq[(LQ_DD + 0) * S + k], d1 = lq[(LQ_DD + 1) * S + k], d2 = lq[(LQ_DD + 2) * S + k];
    const double e03 = lq[(LQ_EE + 0) * S + k], e13 = lq[(LQ_EE + 1) * S + k], e23 = lq[(LQ_EE + 2) * S + k];
    const double e04 = lq[(LQ_EE + 3) * S + k], e14 = lq[(LQ_EE + 4) * S + k];
    const double tv = T * du[0];
    dx[0] += e03 * dx[3] + e04 * dx[4] + tv * d0;
    dx[1] += e13 * dx[3] + e14 * dx[4] + tv * d1;
    dx[2] += e23 * dx[3] + tv * d2;
#pragma unroll
    for (int i = 1; i < 6; ++i) dx[i + 2] += T * du[i];
  }
  if (lane == N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) mydx[i] = dx[i];
  }
  if (lane <= N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dxo[i * S + lane] = mydx[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) duo[i * S + lane] = mydu[i];
  }
  __syncwarp();
}

// Backward sweep for a new state gradient q' (q2[8][S]) with the stored factors (SOC right-hand sides).
__device__ __noinline__ void riccati_resolve(const Prob& pr, const double* lq, const double* q2, double* ric, double mu, int lane) {
  const int S = pr.S, N = pr.N; const double T = pr.T;
  double p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = q2[i * S + N];
#pragma unroll 1
  for (int k = N - 1; k >= 0; --k) {
    double* rk = ric + k * RIC_N;
    const double d0 = lq[(LQ_DD + 0) * S + k], d1 = lq[(LQ_DD + 1) * S + k], d2 = lq[(LQ_DD + 2) * S + k];
    const double e03 = lq[(LQ_EE + 0) * S + k], e13 = lq[(LQ_EE + 1) * S + k], e23 = lq[(LQ_EE + 2) * S + k];
    const double e04 = lq[(LQ_EE + 3) * S + k], e14 = lq[(LQ_EE + 4) * S + k];
    double psi[6], kap[6];
    psi[0] = mu * lq[(LQ_RB + 0) * S + k] + T * (d0 * p[0] + d1 * p[1] + d2 * p[2]);
#pragma unroll
    for (int r = 1; r < 6; ++r) psi[r] = mu * lq[(LQ_RB + r) * S + k] + T * p[r + 2];
    double L[21], idg[6];
#pragma unroll
    for (int e = 0; e < 21; ++e) L[e] = __ldcg(rk + RIC_L + e);
#pragma unroll
    for (int r = 0; r < 6; ++r) { idg[r] = __ldcg(rk + RIC_L + 21 + r); kap[r] = psi[r]; }
    chol6_solve(L, idg, kap);
    double pn[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pn[i] = q2[i * S + k] + p[i];
    pn[3] += e03 * p[0] + e13 * p[1] + e23 * p[2];
    pn[4] += e04 * p[0] + e14 * p[1];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) acc += __ldcg(rk + RIC_K + r * 9 + i) * psi[r];
      pn[i] -= acc;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pn[i];
#pragma unroll
    for (int r = 0; r < 6; ++r) if (lane == r) rk[RIC_K2 + r] = kap[r];
  }
  __syncwarp();
}

}  // namespace nmpc
