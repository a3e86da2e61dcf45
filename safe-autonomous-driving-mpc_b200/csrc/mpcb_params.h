// mpcb_params.h -- mpcb_params (ABI struct) -> DevParams (constants the kernels consume).  Host code, shared by
// libmpcb200.so and by the g++ test build of the device headers (tests/hostbuild).
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "mpcb200.h"
#include "mpcb_solver.cuh"

namespace mpcb {

inline void default_params(mpcb_params* p) {
  memset(p, 0, sizeof(*p));
  p->dt = 0.2; p->N = 5;
  p->u_min[0] = -0.6; p->u_min[1] = -5.0; p->u_max[0] = 0.6; p->u_max[1] = 4.0;
  p->vehicle_radius = 1.0;
  p->w_d = 10.0; p->w_o = 10.0; p->w_v = 5.0; p->w_u1 = 0.5; p->w_u2 = 0.5;
  p->obstacle_safety_distance = 5.0; p->max_time_2_obs = 1.5; p->wheelbase = 2.8; p->lane_width = 3.0;
  p->safe_lane_margin = 0.1;
  p->brake_lookahead = 40.0; p->brake_guess = -2.0;
  p->max_rounds = 10; p->max_segments = 12; p->segment_iters = 10;
  p->rho_lo = 0.1; p->rho_hi = 1e4; p->rho_init = 1.0;
  p->alpha = 1.6;
  p->eps_prim = 1e-9; p->eps_dual = 1e-8; p->eps_infeas = 1e-4;
  p->step_tol = 1e-5; p->feas_tol = 1e-6;
  p->fast_pass = 1;
  p->fast_rho_off = 1e-9; p->fast_rho_on = 1e6;
  p->fast_max_rounds = 6; p->fast_max_segments = 4; p->fast_segment_iters = 2;
  p->coop_pass2 = 1; p->coop_max_batch = 3072;
  p->thread_max_rounds = 5; p->thread_max_segments = 1; p->thread_fail_rounds = 3;
}

// mpcb_params -> constants of the kernels.  pol[0]: robust ladder (x10 per rung with hysteresis, alpha as given,
// infeasibility certificate trusted); pol[1]: two-level policy of the first pass ("off" ~ 0 for inactive rows, "on" = a
// large augmented-Lagrangian weight for active rows, alpha = 1).
inline int derive_params(const mpcb_params& p, DevParams& d) {
  if (p.N != NH) return MPCB_ERR_UNSUPPORTED;
  if (!(p.dt > 0) || p.max_rounds < 1 || p.max_segments < 1 || p.segment_iters < 1) return MPCB_ERR_INVALID;
  if (!(p.rho_lo > 0) || !(p.rho_hi >= p.rho_lo) || !(p.alpha > 0 && p.alpha < 2)) return MPCB_ERR_INVALID;
  if (!(p.fast_rho_off > 0) || !(p.fast_rho_on > p.fast_rho_off) || p.fast_max_rounds < 1 ||
      p.fast_max_segments < 1 || p.fast_segment_iters < 1 || p.thread_max_rounds < 1 || p.thread_max_segments < 1)
    return MPCB_ERR_INVALID;
  memset(&d, 0, sizeof(d));
  d.h = p.dt;
  for (int i = 0; i < 2; ++i) { d.umin[i] = p.u_min[i]; d.umax[i] = p.u_max[i]; }
  d.wd = p.w_d; d.wo = p.w_o; d.wv = p.w_v; d.wu[0] = p.w_u1; d.wu[1] = p.w_u2;
  d.obs_safe = p.obstacle_safety_distance; d.tgap = p.max_time_2_obs;
  d.sld = p.lane_width / 2.0 - p.vehicle_radius - p.safe_lane_margin;   // trajectory_tracking.py:169
  d.alpha_lane[0] = 0.0; d.alpha_lane[1] = p.wheelbase / 2.0; d.alpha_lane[2] = p.wheelbase;
  d.brake_lookahead = p.brake_lookahead; d.brake_guess = p.brake_guess;
  d.max_rounds = p.max_rounds; d.fast_max_rounds = p.fast_max_rounds;
  d.thread_max_rounds = std::min(p.thread_max_rounds, p.fast_max_rounds);
  d.thread_max_segments = std::min(p.thread_max_segments, p.fast_max_segments);
  // robust ladder
  Policy& r = d.pol[0];
  const double fac = 10.0;
  int n = 0;
  double rho = p.rho_lo;
  while (n < MAXRUNG) {
    r.lad[n++] = std::min(rho, p.rho_hi);
    if (rho >= p.rho_hi) break;
    rho *= fac;
  }
  r.n_rung = n;
  r.lad_ratio[0] = 1.0;
  for (int k = 1; k < n; ++k) r.lad_ratio[k] = r.lad[k - 1] / r.lad[k];
  int best = 0;
  for (int k = 0; k < n; ++k)
    if (fabs(log(r.lad[k] / p.rho_init)) < fabs(log(r.lad[best] / p.rho_init))) best = k;
  r.e_init = best;
  r.relax = p.alpha;
  r.hysteresis = 1; r.drop_all = 0;
  r.max_segments = p.max_segments; r.segment_iters = p.segment_iters;
  // two-level policy
  Policy& f = d.pol[1];
  f.lad[0] = p.fast_rho_off; f.lad[1] = p.fast_rho_on;
  f.lad_ratio[0] = 1.0; f.lad_ratio[1] = f.lad[0] / f.lad[1];
  f.n_rung = 2; f.e_init = 0; f.hysteresis = 0; f.drop_all = 1;
  f.relax = 1.0;
  f.max_segments = p.fast_max_segments; f.segment_iters = p.fast_segment_iters;
  d.eps_p = p.eps_prim; d.eps_d = p.eps_dual; d.eps_inf = p.eps_infeas;
  d.step_tol = p.step_tol; d.feas_tol = p.feas_tol;
  d.qp_forcing = 1e-3; d.qp_eps_loose = 1e-4;
  d.max_fail_rounds = 2;
  d.fast_fail_rounds = p.thread_fail_rounds < 0 ? 0 : p.thread_fail_rounds;
#if defined(MPCB_DEV) && !defined(__CUDACC__)   // development / host test builds only; the product reads no environment
  if (getenv("MPCB_FORCING")) d.qp_forcing = atof(getenv("MPCB_FORCING"));
  if (getenv("MPCB_MAXFAIL")) d.max_fail_rounds = atoi(getenv("MPCB_MAXFAIL"));
  if (getenv("MPCB_FASTFAIL")) d.fast_fail_rounds = atoi(getenv("MPCB_FASTFAIL"));
#endif
  const double h = p.dt, floor_ = NRM2_FLOOR;
  for (int j = 1; j <= NH; ++j) {
    double nv = 0, n1 = 0, n2 = 0;
    for (int i = 0; i < j; ++i) {
      const double cs = h * h * (double)(j - 1 - i);
      nv += h * h;
      n1 += cs * cs;
      n2 += (cs + p.max_time_2_obs * h) * (cs + p.max_time_2_obs * h);
    }
    d.inrm_v[j - 1] = 1.0 / std::max(nv, floor_);
    d.inrm_r1[j - 1] = 1.0 / std::max(n1, floor_);
    d.inrm_r2[j - 1] = 1.0 / std::max(n2, floor_);
  }
  return MPCB_OK;
}

}  // namespace mpcb
