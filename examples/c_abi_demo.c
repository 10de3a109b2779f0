/* Plain-C caller of the drop-in boundary (include/nmpc_b200.h): no Python, no torch, no CUDA headers in this file.
 *
 * The first closed-loop steps of Python/T_Trajectory.py -- the script's `while mpc_iter < loop_run` body (same structure as
 * Python/NMPC_TT.py:346-402) -- with the CasADi call replaced by nmpc_solve_host and the shift (NMPC_TT.py:13-30) done on the
 * host in C, exactly where the script does it.  Known answer of the first solve: f* = 248.10109322.
 *
 *   gcc -O2 -Iinclude -o examples/c_abi_demo examples/c_abi_demo.c -Lmpc-implementation_b200/csrc -lnmpc_b200 \
 *       -Wl,-rpath,'$ORIGIN/../mpc-implementation_b200/csrc' -lm
 *   examples/c_abi_demo [steps]            (needs a B200; nmpc_create fails with a message otherwise)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nmpc_b200.h"

#define HORIZON 15
#define NU 6
#define NOBS 3
#define NW (NU * HORIZON)
#define ROWS (5 + NOBS)
#define NG (ROWS * (HORIZON + 1))

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 5;
  const double PI = 3.14159265358979323846, T = 0.2;                   /* T_Trajectory.py: T, N as NMPC_TT.py:57-58 */
  nmpc_spec spec;
  memset(&spec, 0, sizeof spec);
  spec.T = T; spec.N = HORIZON; spec.n_obs = NOBS; spec.w1 = 1.0; spec.w2 = 2.0; spec.vfov = 1.0; spec.hfov = 1.0;
  spec.max_iter = 100; spec.scaling = 1; spec.tol = 1e-8; spec.max_batch = 1; spec.model = NMPC_MODEL_GIMBAL;
  nmpc_handle* h = NULL;
  if (nmpc_create(&spec, 0, &h) != 0) { fprintf(stderr, "nmpc_create: %s\n", nmpc_last_error()); return 2; }

  /* bounds as the script fills args['lbx'] ... args['ubg'] (NMPC_TT.py:269-313) */
  const double ulo[NU] = {14.0, -PI / 30, -PI / 21, -PI / 30, -PI / 30, -PI / 30}, uhi[NU] = {30.0, PI / 30, PI / 21, PI / 30, PI / 30, PI / 30};
  const double glo[5] = {75.0, -0.2618, -PI / 6, -PI / 6, -PI / 2}, ghi[5] = {150.0, 0.2618, PI / 6, PI / 6, PI / 2};
  double lbx[NW], ubx[NW], lbg[NG], ubg[NG];
  for (int k = 0; k < HORIZON; ++k) for (int i = 0; i < NU; ++i) { lbx[NU * k + i] = ulo[i]; ubx[NU * k + i] = uhi[i]; }
  for (int k = 0; k <= HORIZON; ++k) for (int r = 0; r < ROWS; ++r) {
    lbg[ROWS * k + r] = r < 5 ? glo[r] : -INFINITY; ubg[ROWS * k + r] = r < 5 ? ghi[r] : 0.0;
  }
  /* three far-away dummy obstacles, {cx, cy, UAV_r + obs_r} */
  const double obst[NOBS * 3] = {10000.0, 10000.0, 35.0, 10000.0, 10000.0, 35.0, 10000.0, 10000.0, 35.0};

  double x0[8] = {99.0, 150.0, 80.0, 0, 0, 0, 0, 0}, xs[3] = {100.0, 150.0, 0.0};      /* initial UAV / target state */
  double u0[NW] = {0}, p[11], x[NW], f;
  int32_t status, iters;
  for (int mpc_iter = 0; mpc_iter < steps; ++mpc_iter) {
    memcpy(p, x0, sizeof x0); memcpy(p + 8, xs, sizeof xs);                            /* args['p']   NMPC_TT.py:350-353 */
    if (nmpc_solve_host(h, 1, p, u0, lbx, ubx, lbg, ubg, obst, 0, x, &f, NULL, NULL, NULL, &status, &iters) != 0) {   /* :358-365 */
      fprintf(stderr, "nmpc_solve_host: %s\n", nmpc_last_error()); return 3;
    }
    printf("step %d: status %d, %d iterations, f = %.8f, u0 = [%.6f %.6f %.6f %.6f %.6f %.6f]\n", mpc_iter, status, iters, f,
           x[0], x[1], x[2], x[3], x[4], x[5]);
    /* shift_timestep (NMPC_TT.py:13-30): plant Euler step with the first input, warm start shifted, target step */
    const double v = x[0], th = x0[3], ps = x0[4];
    const double rhs[8] = {v * cos(ps) * cos(th), v * sin(ps) * cos(th), v * sin(th), x[1], x[2], x[3], x[4], x[5]};
    for (int i = 0; i < 8; ++i) x0[i] += T * rhs[i];
    memmove(u0, x + NU, sizeof(double) * NU * (HORIZON - 1)); memcpy(u0 + NU * (HORIZON - 1), x + NU * (HORIZON - 1), sizeof(double) * NU);
    const double vt = 13.5, wt = 0.0;                                                  /* T_Trajectory.py:24-57, first 100 steps */
    const double th_t = xs[2];
    xs[0] += T * vt * cos(th_t); xs[1] += T * vt * sin(th_t); xs[2] += T * wt;
  }
  nmpc_destroy(h);
  return 0;
}
