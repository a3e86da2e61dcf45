// mpcb_planner.cuh -- function evaluation of the offline Hermite-Simpson planning NLP (north_star item (c)).
//
// What is computed follows the reference formulation:
//   dynamics                 trajectory_planning.py:49-89   (curvilinear bicycle model, |1 - d k_ref| clamped at 1e-4)
//   Hermite-Simpson defect   trajectory_planning.py:183-208 (same u_k at both ends and at the midpoint)
//   cost                     trajectory_planning.py:128-170
//   inequality closures      trajectory_planning.py:249-347
//   z layout                 trajectory_planning.py:91-126  z = [X (N+1)x5 ; U Nx2 ; S N]
// How it is computed is new: one thread per collocation interval, analytic first and second derivatives
// (the reference finite-differences every closure), results staged in shared memory and written coalesced.
//
// k_ref(s) is the piecewise-linear curvature column of the reference-signal table with linear extrapolation on
// both sides (TrajectoryLoader.interp_k, trajectory_loader.py:69): the GraphHopper spline the reference planner
// used is not in the repository.  k_ref'' = 0 almost everywhere; at a knot the left segment is used
// (searchsorted side='left'), like scipy's interp1d.
//
// simpson_sign = -1 reproduces the committed reference code (x_pred = x_k - dt/6 (...), :198);
// simpson_sign = +1 is the form the committed trajectories/*.json satisfy.
#pragma once
#include "mpcb_device.cuh"

namespace mpcb {

struct PlanParams {
  double dt;
  double w_y, w_s, w_u, w_slack;
  double u_min[2], u_max[2];
  double k_min, k_max, a_max;
  double sigma;            // simpson_sign as a double: defect = x_n - x_k - sigma dt/6 (f_k + 4 f_m + f_n)
  double v_min_c, v_max_c; // used when no per-node arrays are given
  double s_total;
};

// interp_k(s): value and slope of the bracketing segment (no s >= s_max clamp: this is the raw interpolator).
// A lookup is a chain of dependent memory round trips (bucket index -> knots -> rows), and with one or two warps per
// scheduler nobody hides them.  So: the five knots around a first guess and their curvatures are loaded at once and the
// bracketing segment is picked in registers -- the bucket index is at most a knot or two off, and the three points of a
// collocation interval lie within a knot or two of each other (hint: segment of such a neighbour, 0: none).  Two round
// trips without a hint, one with; the walk of seg_index_hint remains as the fallback when the point is outside the
// window.  Same segment either way: the searchsorted-left bracket clipped to [1, K - 1].
MPCB_HD void lookup_kref(const DevTable& T, double s, double& kap, double& dkap, int& hint) {
  const int K = T.K;
  int c = hint;
  if (c <= 0) {
    if (T.lut) {
      double t = (s - T.lut_s0) * T.lut_scale;
      t = t > 0.0 ? t : 0.0;                                // also catches NaN
      t = t < (double)(T.lut_n - 1) ? t : (double)(T.lut_n - 1);
      c = MPCB_LDG(T.lut + (int)t);
    } else {
      c = seg_index(T.s, K, s);
    }
  }
  c = c < 1 ? 1 : (c > K - 1 ? K - 1 : c);
  double sw[5], kw[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    int q = c - 2 + j;
    q = q < 0 ? 0 : (q > K - 1 ? K - 1 : q);
    sw[j] = MPCB_LDG(T.s + q);
    kw[j] = MPCB_LDG(T.y + 4 * q + 2);
  }
  int i = -1;
  double x_lo = 0.0, x_hi = 1.0, y_lo = 0.0, y_hi = 0.0;
#pragma unroll
  for (int j = 4; j >= 1; --j) {                            // candidates c - 1 .. c + 2; the lowest match wins
    const int cand = c - 2 + j;
    const bool ok = cand >= 1 && cand <= K - 1 && (cand == 1 || sw[j - 1] < s) && (cand == K - 1 || sw[j] >= s);
    if (ok) { i = cand; x_lo = sw[j - 1]; x_hi = sw[j]; y_lo = kw[j - 1]; y_hi = kw[j]; }
  }
  if (i < 0) {                                              // outside the window: walk / search, then load
    int h = c;
    i = seg_index_hint(T.s, K, s, h);
    x_lo = MPCB_LDG(T.s + i - 1); x_hi = MPCB_LDG(T.s + i);
    y_lo = MPCB_LDG(T.y + 4 * (i - 1) + 2); y_hi = MPCB_LDG(T.y + 4 * i + 2);
  }
  hint = i;
  const double wl = (s - x_lo) / (x_hi - x_lo), wr = (x_hi - s) / (x_hi - x_lo);
  kap = wl * y_hi + wr * y_lo;
  dkap = (y_hi - y_lo) / (x_hi - x_lo);
}

// Everything about one evaluation point that the derivative formulas reuse.
struct HsPoint {
  double f[5];
  double c, sn, g, kap, dkap, d, k, v;   // cos o, sin o, 1/denom, k_ref, k_ref', state entries
  double gs, gd;                         // d g / d s, d g / d d   (zero when the denominator clamp is active)
  bool clamped;
};

MPCB_HD void hs_point(const DevTable& T, const double (&x)[5], const double (&u)[2], HsPoint& p, int& hint) {
  const double s = x[0], d = x[1], o = x[2], k = x[3], v = x[4];
  lookup_kref(T, s, p.kap, p.dkap, hint);
  double den = 1.0 - d * p.kap;
  p.clamped = fabs(den) < 1e-4;
  if (p.clamped) den = (den != 0.0) ? ((den > 0.0) ? 1e-4 : -1e-4) : 1e-4;       // :74-77
  sincos(o, &p.sn, &p.c);          // one argument reduction for both
  p.g = 1.0 / den;
  p.d = d; p.k = k; p.v = v;
  const double sdot = (v * p.c) / den;
  p.f[0] = sdot;
  p.f[1] = v * p.sn;
  p.f[2] = v * k - sdot * p.kap;
  p.f[3] = u[0];
  p.f[4] = u[1];
  p.gs = p.clamped ? 0.0 : d * p.dkap * p.g * p.g;
  p.gd = p.clamped ? 0.0 : p.kap * p.g * p.g;
}

// Structural zeros, known at compile time (the loops below are fully unrolled, so a masked term costs nothing; a
// multiplication by a stored 0.0 would -- IEEE arithmetic does not let the compiler drop it):
//   F = df/dx          rows s', d', o' only; s' does not depend on k, d' only on o and v
//   M = dx_m/dx_{k,n}  = I/2 +- dt/8 F: F's pattern plus the diagonal
//   W = sum_c w_c Hess(f_c): the block over (s, d, o, v) without (v, v), plus the (k, v) pair
// 332 instead of 775 multiply-adds in the three matrix products of an interval.
__host__ __device__ constexpr bool hs_fnz(int r, int c) { return r == 0 ? (c != 3) : (r == 1 ? (c == 2 || c == 4) : r == 2); }
__host__ __device__ constexpr bool hs_mnz(int r, int c) { return r == c || hs_fnz(r, c); }
__host__ __device__ constexpr bool hs_wnz(int r, int c) {
  return (r == 3 || c == 3) ? ((r == 3 && c == 4) || (r == 4 && c == 3)) : !(r == 4 && c == 4);
}

// F = d f / d x (5x5; rows 3,4 are zero), variables ordered s,d,o,k,v
MPCB_HD void hs_jac(const HsPoint& p, double (&F)[5][5]) {
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) F[r][c] = 0.0;
  const double A = p.v * p.c, Bq = p.v * p.sn;
  F[0][0] = A * p.gs;
  F[0][1] = A * p.gd;
  F[0][2] = -Bq * p.g;
  F[0][4] = p.c * p.g;
  F[1][2] = A;
  F[1][4] = p.sn;
  // f2 = v k - f0 kap
  F[2][0] = -F[0][0] * p.kap - p.f[0] * p.dkap;
  F[2][1] = -F[0][1] * p.kap;
  F[2][2] = -F[0][2] * p.kap;
  F[2][3] = p.v;
  F[2][4] = p.k - F[0][4] * p.kap;
}

// W += sum_c w_c Hess(f_c)  (5x5 symmetric, full storage), c = 0..2 (f3, f4 are linear)
MPCB_HD void hs_hess_acc(const HsPoint& p, const double (&w)[5], double (&W)[5][5]) {
  const double A = p.v * p.c, Bq = p.v * p.sn, g = p.g;
  double gss = 0.0, gsd = 0.0, gdd = 0.0;
  if (!p.clamped) {
    const double g2 = g * g, g3 = g2 * g;
    gss = 2.0 * p.d * p.d * p.dkap * p.dkap * g3;
    gsd = p.dkap * g2 + 2.0 * p.d * p.kap * p.dkap * g3;
    gdd = 2.0 * p.kap * p.kap * g3;
  }
  double H0[5][5];
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) H0[r][c] = 0.0;
  H0[0][0] = A * gss;
  H0[0][1] = H0[1][0] = A * gsd;
  H0[1][1] = A * gdd;
  H0[0][2] = H0[2][0] = -Bq * p.gs;
  H0[1][2] = H0[2][1] = -Bq * p.gd;
  H0[2][2] = -A * g;
  H0[0][4] = H0[4][0] = p.c * p.gs;
  H0[1][4] = H0[4][1] = p.c * p.gd;
  H0[2][4] = H0[4][2] = -p.sn * g;
  const double g0[5] = {A * p.gs, A * p.gd, -Bq * g, 0.0, p.c * g};   // grad f0
  const double a0 = w[0] - p.kap * w[2];
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c)
      if (hs_wnz(r, c) && r != 3 && c != 3) W[r][c] = fma(a0, H0[r][c], W[r][c]);
  // f1 = v sin o
  W[2][2] = fma(w[1], -Bq, W[2][2]);
  W[2][4] = fma(w[1], p.c, W[2][4]);
  W[4][2] = fma(w[1], p.c, W[4][2]);
  // f2: -kap' (e_s grad f0' + grad f0 e_s') + (e_k e_v' + e_v e_k')
  const double t = -w[2] * p.dkap;
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    if (c == 3) continue;                       // grad f0 has no k component
    W[0][c] = fma(t, g0[c], W[0][c]);
    W[c][0] = fma(t, g0[c], W[c][0]);
  }
  W[3][4] += w[2];
  W[4][3] += w[2];
}

// One collocation interval.  lam may be null when hess is not wanted.
//   defect[5], jac[5][12] (columns: x_k(5), x_{k+1}(5), u_k(2)), hess[12][12] of lam . defect
//   HESS_TRI: hess receives only the 55 entries of the lower triangle of the 10 x 10 state block, packed row by row
//   (entry (i, j), j <= i, at i (i + 1) / 2 + j); the controls enter the defect linearly, so rows and columns 10, 11 of
//   the 12 x 12 block are zero and the block is symmetric -- the kernel expands it while writing it out.
template <bool WANT_JAC, bool WANT_HESS, bool HESS_TRI = false>
MPCB_HD void hs_interval(const DevTable& T, const PlanParams& P, const double (&xk)[5], const double (&xn)[5],
                         const double (&u)[2], const double* lam, double* defect, double* jac, double* hess) {
  const double dt = P.dt, sg = P.sigma;
  HsPoint pk, pn, pm;
  int hint = 0, hint_n = 0;            // the end points look their segments up independently (in parallel), the
  hs_point(T, xk, u, pk, hint);        // midpoint starts from the first one's
  hs_point(T, xn, u, pn, hint_n);
  double xm[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) xm[c] = 0.5 * (xk[c] + xn[c]) + (dt / 8.0) * (pk.f[c] - pn.f[c]);     // :195
  hs_point(T, xm, u, pm, hint);
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const double simpson = (dt / 6.0) * (pk.f[c] + 4 * pm.f[c] + pn.f[c]);
    const double x_pred = (sg < 0.0) ? (xk[c] - simpson) : (xk[c] + simpson);                       // :198 / C7
    defect[c] = xn[c] - x_pred;                                                                     // :200
  }
  if (!WANT_JAC && !WANT_HESS) return;
  double Fk[5][5], Fn[5][5], Fm[5][5];
  hs_jac(pk, Fk);
  hs_jac(pn, Fn);
  hs_jac(pm, Fm);
  double Mk[5][5], Mn[5][5];   // d x_m / d x_k, d x_m / d x_n
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      if (hs_fnz(r, c)) {
        Mk[r][c] = ((r == c) ? 0.5 : 0.0) + (dt / 8.0) * Fk[r][c];
        Mn[r][c] = ((r == c) ? 0.5 : 0.0) - (dt / 8.0) * Fn[r][c];
      } else {
        Mk[r][c] = (r == c) ? 0.5 : 0.0;
        Mn[r][c] = (r == c) ? 0.5 : 0.0;
      }
    }
  const double cf = -sg * dt / 6.0;
  if (WANT_JAC) {
#pragma unroll
    for (int r = 0; r < 5; ++r) {
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        double fk = 0.0, fn = 0.0;
#pragma unroll
        for (int m = 0; m < 5; ++m)
          if (hs_fnz(r, m) && hs_mnz(m, c)) { fk = fma(Fm[r][m], Mk[m][c], fk); fn = fma(Fm[r][m], Mn[m][c], fn); }
        jac[r * 12 + c] = ((r == c) ? -1.0 : 0.0) + cf * (Fk[r][c] + 4.0 * fk);
        jac[r * 12 + 5 + c] = ((r == c) ? 1.0 : 0.0) + cf * (Fn[r][c] + 4.0 * fn);
      }
      // u enters f additively (rows 3, 4) and cancels in x_m: d defect / d u = -sigma dt B
      jac[r * 12 + 10] = (r == 3) ? 6.0 * cf : 0.0;
      jac[r * 12 + 11] = (r == 4) ? 6.0 * cf : 0.0;
    }
  }
  if (WANT_HESS) {
    double l[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) l[c] = lam[c];
    double gm[5];   // grad_x (lam . f)(x_m) = Fm' lam
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      double a = 0.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) a = fma(Fm[r][c], l[r], a);
      gm[c] = a;
    }
    double Wm[5][5], Wk[5][5], Wn[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int c = 0; c < 5; ++c) { Wm[r][c] = 0.0; Wk[r][c] = 0.0; Wn[r][c] = 0.0; }
    hs_hess_acc(pm, l, Wm);
    double wk[5], wn[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) { wk[c] = l[c] + 0.5 * dt * gm[c]; wn[c] = l[c] - 0.5 * dt * gm[c]; }
    hs_hess_acc(pk, wk, Wk);
    hs_hess_acc(pn, wn, Wn);
    // T = Wm [Mk Mn]  (5x10), then H = cf (blockdiag(Wk, Wn) + 4 [Mk Mn]' T)
    double Tm[5][10];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int m = 0; m < 5; ++m)
          if (hs_wnz(r, m) && hs_mnz(m, c)) { a = fma(Wm[r][m], Mk[m][c], a); b = fma(Wm[r][m], Mn[m][c], b); }
        Tm[r][c] = a; Tm[r][5 + c] = b;
      }
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        if (i >= 10 || j >= 10) { if (!HESS_TRI) hess[i * 12 + j] = 0.0; continue; }
        if (j > i) continue;   // lower triangle first
        double a = 0.0;
#pragma unroll
        for (int m = 0; m < 5; ++m) {
          if (!hs_mnz(m, i < 5 ? i : i - 5)) continue;
          const double mi = (i < 5) ? Mk[m][i] : Mn[m][i - 5];
          a = fma(mi, Tm[m][j], a);
        }
        double direct = 0.0;
        if (i < 5 && j < 5) direct = Wk[i][j];
        if (i >= 5 && j >= 5) direct = Wn[i - 5][j - 5];
        const double val = cf * (direct + 4.0 * a);
        if (HESS_TRI) { hess[i * (i + 1) / 2 + j] = val; continue; }
        hess[i * 12 + j] = val;
        hess[j * 12 + i] = val;
      }
  }
}

#if defined(__CUDACC__)
// The same interval evaluated by TWO lanes (lane pair p = 0 / 1 of a warp, partner = lane ^ 1): what the GPU kernel runs.
// Lane p owns the end point x_k (p = 0) resp. x_{k+1} (p = 1): its dynamics, Jacobian F_p, M_p = dx_m/dx_p, the weighted
// Hessian W_p and T_p = W_m M_p; the midpoint (which needs both ends' dynamics: three doubles cross the pair) is
// evaluated by both.  Both lanes run ONE instruction stream on mirrored data:
//   Jacobian   lane p writes the five columns of its own state
//   Hessian    lane p writes the triangle of its own diagonal block, cf (W_p + 4 M_p' T_p), and half of the cross block
//              4 cf M_n' W_m M_k: with C_p[a][b] = sum_m M_partner[m][a] T_p[m][b], lane 0 holds H_nk[a][b] and lane 1
//              H_nk[b][a], so each takes the entries a <= b (lane 0 keeps the diagonal); M's 13 structural nonzeros
//              are what the lanes exchange.
// 170 multiply-adds per lane in the matrix products instead of 332 in one thread, about half the registers, and twice
// the warps per SM.  slot: the interval's staging area (defect 5, Jacobian 60 at +5, Hessian triangle 55 at +65).
// pair_mask: the lanes of the warp that are inside hs_interval_pair (whole pairs).
__device__ __forceinline__ double hs_xchg(unsigned mask, double v) { return __shfl_xor_sync(mask, v, 1); }

template <bool WANT_JAC, bool WANT_HESS>
__device__ __forceinline__ void hs_interval_pair(const DevTable& T, const PlanParams& P, const double (&xk)[5],
                                                 const double (&xn)[5], const double (&u)[2], const double* lam, int p,
                                                 unsigned pair_mask, double* slot) {
  const double dt = P.dt, sg = P.sigma;
  const double sgn = p ? -1.0 : 1.0;
  double l[5] = {0.0, 0.0, 0.0, 0.0, 0.0};      // multipliers: loaded up front, consumed a thousand instructions later
  if (WANT_HESS) {
#pragma unroll
    for (int c = 0; c < 5; ++c) l[c] = MPCB_LDG(lam + c);
  }
  double xo[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) xo[c] = p ? xn[c] : xk[c];
  HsPoint po, pm;
  int hint = 0;
  hs_point(T, xo, u, po, hint);
  double fk[5], fn[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const double other = (c < 3) ? hs_xchg(pair_mask, po.f[c]) : u[c - 3];
    fk[c] = p ? other : po.f[c];
    fn[c] = p ? po.f[c] : other;
  }
  double xm[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) xm[c] = 0.5 * (xk[c] + xn[c]) + (dt / 8.0) * (fk[c] - fn[c]);        // :195
  hs_point(T, xm, u, pm, hint);
  if (p == 0) {
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const double simpson = (dt / 6.0) * (fk[c] + 4 * pm.f[c] + fn[c]);
      const double x_pred = (sg < 0.0) ? (xk[c] - simpson) : (xk[c] + simpson);                     // :198 / C7
      slot[c] = xn[c] - x_pred;                                                                     // :200
    }
  }
  if (!WANT_JAC && !WANT_HESS) return;
  double Fo[5][5], Fm[5][5];
  hs_jac(po, Fo);
  hs_jac(pm, Fm);
  double Mo[5][5];            // d x_m / d x_own
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      if (hs_fnz(r, c)) Mo[r][c] = ((r == c) ? 0.5 : 0.0) + sgn * ((dt / 8.0) * Fo[r][c]);
      else Mo[r][c] = (r == c) ? 0.5 : 0.0;
    }
  const double cf = -sg * dt / 6.0;
  if (WANT_JAC) {
    double* jac = slot + 5;
    const int col0 = 5 * p;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        double fo = 0.0;
#pragma unroll
        for (int m = 0; m < 5; ++m)
          if (hs_fnz(r, m) && hs_mnz(m, c)) fo = fma(Fm[r][m], Mo[m][c], fo);
        jac[r * 12 + col0 + c] = ((r == c) ? -sgn : 0.0) + cf * (Fo[r][c] + 4.0 * fo);
      }
      if (p == 0) {             // u enters f additively (rows 3, 4) and cancels in x_m: d defect / d u = -sigma dt B
        jac[r * 12 + 10] = (r == 3) ? 6.0 * cf : 0.0;
        jac[r * 12 + 11] = (r == 4) ? 6.0 * cf : 0.0;
      }
    }
  }
  if (WANT_HESS) {
    double* tri_out = slot + 65;
    double gm[5];   // grad_x (lam . f)(x_m) = Fm' lam
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      double a = 0.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) a = fma(Fm[r][c], l[r], a);
      gm[c] = a;
    }
    double Wm[5][5], Wo[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int c = 0; c < 5; ++c) { Wm[r][c] = 0.0; Wo[r][c] = 0.0; }
    hs_hess_acc(pm, l, Wm);
    double wo[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) wo[c] = l[c] + sgn * (0.5 * dt * gm[c]);
    hs_hess_acc(po, wo, Wo);
    double To[5][5];          // W_m M_own
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        double a = 0.0;
#pragma unroll
        for (int m = 0; m < 5; ++m)
          if (hs_wnz(r, m) && hs_mnz(m, c)) a = fma(Wm[r][m], Mo[m][c], a);
        To[r][c] = a;
      }
    double Mx[5][5];          // the partner's M (its structural nonzeros cross the pair)
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int c = 0; c < 5; ++c) Mx[r][c] = hs_mnz(r, c) ? hs_xchg(pair_mask, Mo[r][c]) : 0.0;
    // own diagonal block: rows / columns 5 p + (0..4)
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double a = 0.0;
#pragma unroll
        for (int m = 0; m < 5; ++m)
          if (hs_mnz(m, i)) a = fma(Mo[m][i], To[m][j], a);
        const double val = cf * (Wo[i][j] + 4.0 * a);
        tri_out[p ? ((5 + i) * (6 + i) / 2 + 5 + j) : (i * (i + 1) / 2 + j)] = val;
      }
    // cross block, entries a <= b of C[a][b] = sum_m Mx[m][a] To[m][b]: lane 0 -> H(5 + a, b), lane 1 -> H(5 + b, a)
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int b = a; b < 5; ++b) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < 5; ++m)
          if (hs_mnz(m, a)) acc = fma(Mx[m][a], To[m][b], acc);
        const double val = cf * (0.0 + 4.0 * acc);
        if (p == 0) tri_out[(5 + a) * (6 + a) / 2 + b] = val;
        else if (a != b) tri_out[(5 + b) * (6 + b) / 2 + a] = val;
      }
  }
}
#endif

// Node rows (trajectory_planning.py:249-307), 6 per node k = 0..N:
//   (v+slack) - v_min, v_max - (v+slack), a_max - k v^2, a_max + k v^2, k - k_min, k_max - k      (slack = 0 at k = N)
MPCB_HD void hs_node_rows(const PlanParams& P, const double (&x)[5], double slack, double vmin, double vmax,
                          double (&out)[6]) {
  const double k = x[3], v = x[4];
  out[0] = (v + slack) - vmin;
  out[1] = vmax - (v + slack);
  out[2] = P.a_max - (k * (v * v));
  out[3] = P.a_max + (k * (v * v));
  out[4] = k - P.k_min;
  out[5] = P.k_max - k;
}

// Control rows (:310-347), 5 per interval: u1 - u1min, u1max - u1, u2 - u2min, u2max - u2, slack
MPCB_HD void hs_ctrl_rows(const PlanParams& P, const double (&u)[2], double slack, double (&out)[5]) {
  out[0] = u[0] - P.u_min[0];
  out[1] = P.u_max[0] - u[0];
  out[2] = u[1] - P.u_min[1];
  out[3] = P.u_max[1] - u[1];
  out[4] = slack;
}

// Stage cost of interval k (:143-169) and its gradient w.r.t. (s_k, d_k, o_k, u1_k, u2_k, slack_k)
MPCB_HD double hs_stage_cost(const PlanParams& P, const double (&x)[5], const double (&u)[2], double slack, double s0,
                             double* grad6) {
  const double denom = fmax(1.0, P.s_total - s0);                       // :157
  const double e = (P.s_total - x[0]) / denom;
  const double term1 = P.w_y * (x[1] * x[1] + x[2] * x[2]);
  const double term2 = P.w_s * (e * e);
  const double term3 = P.w_u * (u[0] * u[0] + u[1] * u[1]);
  const double term4 = P.w_slack * (slack * slack);
  if (grad6) {
    grad6[0] = -2.0 * P.w_s * e / denom;
    grad6[1] = 2.0 * P.w_y * x[1];
    grad6[2] = 2.0 * P.w_y * x[2];
    grad6[3] = 2.0 * P.w_u * u[0];
    grad6[4] = 2.0 * P.w_u * u[1];
    grad6[5] = 2.0 * P.w_slack * slack;
  }
  return term1 + term2 + term3 + term4;
}

}  // namespace mpcb
