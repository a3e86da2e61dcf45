// hostlib.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the device headers of libmpcb200.so with g++ so that the
// algorithms (planner derivative formulas, tracking solver) can be validated on the CPU against the oracle
// before / without a GPU.  Nothing in the product package loads this library; the product has no CPU path.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#define __forceinline__ inline
#define __restrict__
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
#define __all_sync(mask, pred) (pred)

#include "mpcb200.h"
#include "mpcb_planner.cuh"
#ifdef MPCB_HOST_SOLVER
#include "mpcb_solver.cuh"
#endif

using namespace mpcb;

static DevTable make_table(const double* s, const double* y, const double* u, int K, int Ku, double s_max,
                           const double* last4) {
  DevTable T;
  T.s = s; T.y = y; T.u = u; T.K = K; T.Ku = Ku; T.s_max = s_max;
  T.lut = nullptr; T.lut_n = 0; T.lut_s0 = 0.0; T.lut_scale = 0.0; T.sinv = nullptr;
  for (int c = 0; c < 4; ++c) T.last[c] = last4[c];
  return T;
}

extern "C" {

// n intervals given as separate arrays; outputs like mpcb_hs_eval (jac/hess may be null)
int host_hs_eval(const double* s, const double* y, int K, double dt, int simpson_sign, int n, const double* xk,
                 const double* xn, const double* u, const double* lam, double* defect, double* jac, double* hess) {
  const double last4[4] = {0, 0, 0, 0};
  DevTable T = make_table(s, y, nullptr, K, K - 1, s[K - 1], last4);
  PlanParams P;
  memset(&P, 0, sizeof(P));
  P.dt = dt; P.sigma = (double)simpson_sign;
  for (int i = 0; i < n; ++i) {
    double a[5], b[5], uu[2], d[5], J[60], H[144];
    for (int c = 0; c < 5; ++c) { a[c] = xk[5 * i + c]; b[c] = xn[5 * i + c]; }
    uu[0] = u[2 * i]; uu[1] = u[2 * i + 1];
    if (hess) hs_interval<true, true>(T, P, a, b, uu, lam + 5 * i, d, J, H);
    else if (jac) hs_interval<true, false>(T, P, a, b, uu, nullptr, d, J, H);
    else hs_interval<false, false>(T, P, a, b, uu, nullptr, d, J, H);
    memcpy(defect + 5 * i, d, sizeof(d));
    if (jac) memcpy(jac + 60 * i, J, sizeof(J));
    if (hess) memcpy(hess + 144 * i, H, sizeof(H));
  }
  return 0;
}

}  // extern "C"

#ifdef MPCB_HOST_SOLVER
#include "mpcb_params.h"

static const double* g_hot_U = nullptr;   // development: [B][10] controls the FIRST pass starts from (instead of the warm start)
static int* g_pass2_stats = nullptr;   // development: [B][3] rounds, iterations, status of the robust pass alone

extern "C" {

void host_set_pass2_stats(int* p) { g_pass2_stats = p; }
void host_set_hot_start(const double* U) { g_hot_U = U; }

// The tracking solver (solve_one, the code mpcb_solve_kernel runs per thread) on the CPU, one problem at a time,
// with the same two-pass logic as launch_solve.  `p` may be null (defaults).  Outputs like mpcb_solve_batch;
// pass_out[b] = 1 when the first pass certified the problem, 2 when the robust pass produced the answer.
int host_solve_batch(const double* s, const double* y, const double* u, int K, int Ku, double s_max,
                     const double* last4, const mpcb_params* p, int B, const double* x0, const double* obs_sv,
                     const int* n_obs, double* U_out, int* status_out, int* iters_out, double* obj_out,
                     int* pass_out) {
  DevTable T = make_table(s, y, u, K, Ku, s_max, last4);
  mpcb_params pp;
  if (p) pp = *p; else default_params(&pp);
  DevParams P;
  int rc = derive_params(pp, P);
  if (rc != 0) return rc;
  const bool carry = getenv("MPCB_CARRY") && atoi(getenv("MPCB_CARRY")) != 0;   // development: pass 2 continues from pass 1's controls
  for (int b = 0; b < B; ++b) {
    int rounds = 0, iters = 0;
    double U1[NV];
    bool have1 = false;
    for (int pass = pp.fast_pass ? 1 : 2; pass <= 2; ++pass) {
      Problem pb;
      for (int c = 0; c < 5; ++c) pb.x0[c] = x0[5 * b + c];
      for (int k = 0; k < 2; ++k) { pb.obs[k][0] = obs_sv[4 * b + 2 * k]; pb.obs[k][1] = obs_sv[4 * b + 2 * k + 1]; }
      pb.n_obs = std::min(std::max(n_obs[b], 0), 2);
      typedef Store<1, 0u> HostStore;
      double buf[HostStore::LOCAL];
      HostStore st(nullptr, buf);
      SolveOut so = (pass == 1) ? solve_one<true>(T, P, pb, st, true, g_hot_U ? g_hot_U + 10 * b : nullptr)
                                : solve_one<false>(T, P, pb, st, true, (carry && have1) ? U1 : nullptr);
      rounds += so.rounds; iters += so.iters;
      if (pass == 2 && g_pass2_stats) { g_pass2_stats[3 * b] = so.rounds; g_pass2_stats[3 * b + 1] = so.iters; g_pass2_stats[3 * b + 2] = so.status; }
      if (pass == 1) { for (int i = 0; i < NV; ++i) U1[i] = pb.U[i]; have1 = true; }
      if (pass == 1 && so.status == MPCB_MAXITER) continue;
      double X[NH + 1][5];
      double cost;
      rollout_values(T, P, pb.x0, pb.U, X, cost);
      double cmin = BIG;
      for (int j = 1; j <= NH; ++j) {
        double rows[9];
        const int nr = constraint_rows(P, X[j], j, pb.obs, pb.n_obs, rows);
        for (int r = 0; r < nr; ++r) cmin = fmin(cmin, rows[r]);
      }
      int status = so.status;
      if (cmin < -P.feas_tol) {
        if (pass == 1 && !so.const_infeasible) continue;
        status = MPCB_INFEASIBLE;
      }
      for (int i = 0; i < NV; ++i) U_out[10 * b + i] = pb.U[i];
      if (status_out) status_out[b] = status;
      if (iters_out) { iters_out[2 * b] = rounds; iters_out[2 * b + 1] = iters; }
      if (obj_out) obj_out[b] = cost;
      if (pass_out) pass_out[b] = pass;
      break;
    }
  }
  return 0;
}

}  // extern "C"
#endif
