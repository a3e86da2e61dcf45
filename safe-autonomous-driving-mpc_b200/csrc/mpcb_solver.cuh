// mpcb_solver.cuh -- per-thread Gauss-Newton SQP with an OSQP-style ADMM QP solver.
//
// QP of one linearisation round, in the absolute variable x = U+ (SURVEY.md A.1 for the formulation):
//     min 1/2 x'Hx + q'x   s.t.  lo <= A x <= hi ,   H = I*2w_u + 2 J'WJ  (Gauss-Newton),  q = g - H U
// Rows of A (M = 47):
//     0..9    box           x_i in [u_min, u_max]                                (trajectory_tracking.py:249)
//     10..21  lane          (D_j + alpha_a O_j) x, j = 2..5, a = 0..2            (:171-189; step 1 is constant in U)
//     22..26  speed         v_j - v0 = h sum_{i<j} b_i >= -v0                    (:207)
//     27+10k+2(j-1)+{0,1}   obstacle k, step j: R1 = S_j <= base - safe,  R2 = S_j + tgap (v_j - v0) <= base - tgap v0
//                           with S_j = s_j - s0 - j h v0 (the max(.,.) of :201 split into two affine rows, A.1)
// ADMM with sigma = 0 and relaxation alpha in single-vector form (v = z_relaxed + y/rho; z = clip(v), y = rho (v - z)):
//     x  = K^-1 (A' rho (2 clip(v) - v) - q),  K = H + A' diag(rho) A ;   v += alpha (A x - clip(v))
// rho_i = lad[e_i] / max(|a_i|^2, floor); every segment (segment_iters iterations) a row that stayed active
// moves one rung up the ladder, a row that stayed inactive one rung down, K is refactored, residuals and the
// OSQP primal-infeasibility certificate are evaluated.
#pragma once
#include <type_traits>

#include "mpcb200.h"
#include "mpcb_device.cuh"

// Loop control is uniform over the whole CTA (all threads reach every barrier): the CTA walks through the
// straight-line solver code together, so one instruction stream per CTA goes through the instruction caches.
#if defined(__CUDA_ARCH__)
#define MPCB_ALL(pred) (__syncthreads_and(pred) != 0)
#else
#define MPCB_ALL(pred) (pred)
#endif

namespace mpcb {

constexpr int M_LANE = 12;
constexpr int ROW_LANE = 10, ROW_V = 22, ROW_OBS = 27, M_ROWS = 47;
constexpr double NRM2_FLOOR = 1e-2;

struct Rungs {  // 4-bit ladder index per row
  unsigned int w[6];
  __device__ __forceinline__ int get(int r) const { return (w[r >> 3] >> ((r & 7) * 4)) & 15; }
  __device__ __forceinline__ void set(int r, int e) {
    w[r >> 3] = (w[r >> 3] & ~(15u << ((r & 7) * 4))) | ((unsigned)e << ((r & 7) * 4));
  }
};

struct Problem {
  // inputs
  double x0[5];
  double obs[2][2];
  int n_obs;
  // SQP state
  double U[NV];
  double x[NV];
  // QP data of the current round
  double H[NTRI];
  double Kinv[NTRI];
  double q[NV];
  double D[4][NV], O[4][NV];   // d(d_j)/dU, d(o_j)/dU for j = 2..5 (only the first 2(j-1) entries are non-zero)
  double lane_c[M_LANE];       // row value offset: (d_j + alpha o_j)(U) - a.U
  double lane_inrm[M_LANE];
  double base[2][NH];          // s_obs_k + v_obs_k j h - s0 - j h v0
  // ADMM state
  double vb[NV], vl[M_LANE], vv[NH], vo[2][NH][2];
  Rungs E;
  unsigned long long act_prev;
};

// ---- bounds per row type ------------------------------------------------------------------------
__device__ __forceinline__ double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// ------------------------------------------------------------------------------------------------
// Linearisation: rollout with forward sensitivities at pb.U; fills H, q, D, O, lane_c, lane_inrm.
// Returns the worst violation among rows that U cannot influence (step-1 lane rows, step-1 "gap - safe"
// row) through const_viol.
// ------------------------------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ void rank1(double (&H)[NTRI], const double (&a)[NV], double w) {
  // H += w * a a'  restricted to the leading L x L block
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const double wa = w * a[i];
#pragma unroll
    for (int j = 0; j <= i; ++j) H[tri(i, j)] = fma(wa, a[j], H[tri(i, j)]);
  }
}

template <int J>  // residual rows of step J (1..5): accumulate H and g
__device__ __forceinline__ void accumulate_step(const DevParams& P, Problem& pb, double (&g)[NV],
                                                const double (&dD)[NV], const double (&dO)[NV],
                                                const double (&Xj)[5], const double (&val)[4],
                                                const double (&slope)[4]) {
  constexpr int L = 2 * (J - 1);  // support of dD, dO, ds_J
  const double h = P.h;
  double Jd[NV], Jo[NV], Jv[NV];
#pragma unroll
  for (int c = 0; c < NV; ++c) { Jd[c] = 0.0; Jo[c] = 0.0; Jv[c] = 0.0; }
#pragma unroll
  for (int i = 0; i < J; ++i) {
    const double ds = (i < J - 1) ? h * h * (double)(J - 1 - i) : 0.0;  // d s_J / d b_i
    Jv[2 * i + 1] = h - slope[3] * ds;
    if (i < J - 1) {
      Jd[2 * i] = dD[2 * i];
      Jo[2 * i] = dO[2 * i];
      Jd[2 * i + 1] = dD[2 * i + 1] - slope[0] * ds;
      Jo[2 * i + 1] = dO[2 * i + 1] - slope[1] * ds;
    }
  }
  const double rd = Xj[1] - val[0], ro = Xj[2] - val[1], rv = Xj[4] - val[3];
  if (L > 0) {
    rank1<L>(pb.H, Jd, 2.0 * P.wd);
    rank1<L>(pb.H, Jo, 2.0 * P.wo);
#pragma unroll
    for (int c = 0; c < L; ++c) g[c] += 2.0 * (P.wd * rd * Jd[c] + P.wo * ro * Jo[c]);
  }
  // Jv lives on the b entries only
#pragma unroll
  for (int i = 0; i < J; ++i) {
    const double wa = 2.0 * P.wv * Jv[2 * i + 1];
#pragma unroll
    for (int k = 0; k <= i; ++k) pb.H[tri(2 * i + 1, 2 * k + 1)] = fma(wa, Jv[2 * k + 1], pb.H[tri(2 * i + 1, 2 * k + 1)]);
    g[2 * i + 1] += 2.0 * P.wv * rv * Jv[2 * i + 1];
  }
}

template <int J>  // lane rows of step J (2..5)
__device__ __forceinline__ void lane_rows(const DevParams& P, Problem& pb, const double (&dD)[NV],
                                          const double (&dO)[NV], const double (&Xj)[5]) {
  constexpr int L = 2 * (J - 1);
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    pb.D[J - 2][c] = (c < L) ? dD[c] : 0.0;
    pb.O[J - 2][c] = (c < L) ? dO[c] : 0.0;
  }
  double dU = 0.0, oU = 0.0, dd = 0.0, dox = 0.0, oo = 0.0;
#pragma unroll
  for (int c = 0; c < L; ++c) {
    dU = fma(dD[c], pb.U[c], dU);
    oU = fma(dO[c], pb.U[c], oU);
    dd = fma(dD[c], dD[c], dd);
    dox = fma(dD[c], dO[c], dox);
    oo = fma(dO[c], dO[c], oo);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double al = P.alpha_lane[a];
    pb.lane_c[3 * (J - 2) + a] = (Xj[1] + al * Xj[2]) - (dU + al * oU);
    const double n2 = dd + 2.0 * al * dox + al * al * oo;
    pb.lane_inrm[3 * (J - 2) + a] = 1.0 / fmax(n2, NRM2_FLOOR);
  }
}

__device__ __forceinline__ void linearise(const DevTable& T, const DevParams& P, Problem& pb, double& const_viol) {
  const double h = P.h;
#pragma unroll
  for (int i = 0; i < NTRI; ++i) pb.H[i] = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i) pb.H[tri(i, i)] = 2.0 * P.wu[i & 1];
  double g[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) g[i] = 2.0 * P.wu[i & 1] * pb.U[i];

  double X[5] = {pb.x0[0], pb.x0[1], pb.x0[2], pb.x0[3], pb.x0[4]};
  double dD[NV], dO[NV];
#pragma unroll
  for (int c = 0; c < NV; ++c) { dD[c] = 0.0; dO[c] = 0.0; }
  double val[4] = {0.0, 0.0, 0.0, 0.0}, slope[4] = {0.0, 0.0, 0.0, 0.0};
  double cv = 0.0;

  auto advance = [&](auto jtag) {
    constexpr int j = decltype(jtag)::value;  // step j -> j+1
    const double s = X[0], d = X[1], o = X[2], k = X[3], v = X[4];
    const double kk = k - val[2];
    // sensitivities first (they use the old state)
    double nD[NV], nO[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) { nD[c] = dD[c]; nO[c] = dO[c]; }
#pragma unroll
    for (int i = 0; i < j; ++i) {
      const double ds = (i < j - 1) ? h * h * (double)(j - 1 - i) : 0.0;
      // a_i :  dk_j = h, dv_j = 0
      nD[2 * i] = dD[2 * i] + h * (v * dO[2 * i]);
      nO[2 * i] = dO[2 * i] + h * (v * h);
      // b_i :  dv_j = h, ds_j = ds
      nD[2 * i + 1] = dD[2 * i + 1] + h * (h * o + v * dO[2 * i + 1]);
      nO[2 * i + 1] = dO[2 * i + 1] + h * (h * kk - v * (slope[2] * ds));
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) { dD[c] = nD[c]; dO[c] = nO[c]; }
    X[0] = s + h * v;
    X[1] = d + h * (v * o);
    X[2] = o + h * (v * kk);
    X[3] = k + h * pb.U[2 * j];
    X[4] = v + h * pb.U[2 * j + 1];
  };

  lookup_state(T, X[0], val, slope);
  advance(std::integral_constant<int, 0>{});
  // step 1: rows constant in U
  lookup_state(T, X[0], val, slope);
  accumulate_step<1>(P, pb, g, dD, dO, X, val, slope);
#pragma unroll
  for (int a = 0; a < 3; ++a) cv = fmax(cv, fabs(X[1] + P.alpha_lane[a] * X[2]) - P.sld);
  advance(std::integral_constant<int, 1>{});
  lookup_state(T, X[0], val, slope);
  accumulate_step<2>(P, pb, g, dD, dO, X, val, slope);
  lane_rows<2>(P, pb, dD, dO, X);
  advance(std::integral_constant<int, 2>{});
  lookup_state(T, X[0], val, slope);
  accumulate_step<3>(P, pb, g, dD, dO, X, val, slope);
  lane_rows<3>(P, pb, dD, dO, X);
  advance(std::integral_constant<int, 3>{});
  lookup_state(T, X[0], val, slope);
  accumulate_step<4>(P, pb, g, dD, dO, X, val, slope);
  lane_rows<4>(P, pb, dD, dO, X);
  advance(std::integral_constant<int, 4>{});
  lookup_state(T, X[0], val, slope);
  accumulate_step<5>(P, pb, g, dD, dO, X, val, slope);
  lane_rows<5>(P, pb, dD, dO, X);

  // q = g - H U
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double acc = g[i];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc = fma(-pb.H[i >= j ? tri(i, j) : tri(j, i)], pb.U[j], acc);
    pb.q[i] = acc;
  }
  // constant obstacle row (step 1, "gap - safe"): base_1 - obs_safe >= 0
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (k < pb.n_obs) cv = fmax(cv, P.obs_safe - pb.base[k][0]);
  const_viol = cv;
}

// ------------------------------------------------------------------------------------------------
// K = H + A' diag(rho) A  ->  Kinv (explicit inverse through Cholesky), all in packed lower triangles.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double row_rho(const DevParams& P, const Rungs& E, int r, double inrm) {
  return P.lad[E.get(r)] * inrm;
}

__device__ __forceinline__ void factor(const DevParams& P, Problem& pb) {
  const double h = P.h;
  double K[NTRI];
#pragma unroll
  for (int i = 0; i < NTRI; ++i) K[i] = pb.H[i];
#pragma unroll
  for (int i = 0; i < NV; ++i) K[tri(i, i)] += row_rho(P, pb.E, i, 1.0);
  // lane rows of step j:  sum_a rho_a (D + al_a O)(D + al_a O)' = D (r0 D + r1 O)' + O (r1 D + r2 O)'
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int L = 2 * (jj + 1);
    double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double rho = row_rho(P, pb.E, ROW_LANE + 3 * jj + a, pb.lane_inrm[3 * jj + a]);
      const double al = P.alpha_lane[a];
      r0 += rho; r1 = fma(rho, al, r1); r2 = fma(rho, al * al, r2);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i < L) {
        const double p = r0 * pb.D[jj][i] + r1 * pb.O[jj][i];
        const double t = r1 * pb.D[jj][i] + r2 * pb.O[jj][i];
#pragma unroll
        for (int j = 0; j <= i; ++j) K[tri(i, j)] = fma(p, pb.D[jj][j], fma(t, pb.O[jj][j], K[tri(i, j)]));
      }
    }
  }
  // constant-coefficient rows act on the b entries only: Kb[i][k] += sum_r rho_r c_r[i] c_r[k]
  double Kb[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) Kb[i] = 0.0;
#pragma unroll
  for (int j = 1; j <= NH; ++j) {
    const double rv = row_rho(P, pb.E, ROW_V + j - 1, P.inrm_v[j - 1]);
    double r1 = 0.0, r2 = 0.0;   // summed over obstacles (same coefficient vectors)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k < pb.n_obs) {
        if (j > 1) r1 += row_rho(P, pb.E, ROW_OBS + 10 * k + 2 * (j - 1), P.inrm_r1[j - 1]);
        r2 += row_rho(P, pb.E, ROW_OBS + 10 * k + 2 * (j - 1) + 1, P.inrm_r2[j - 1]);
      }
    }
#pragma unroll
    for (int i = 0; i < j; ++i) {
      const double cs_i = h * h * (double)(j - 1 - i);   // S_j coefficient
      const double c2_i = cs_i + P.tgap * h;              // R2 coefficient
#pragma unroll
      for (int k = 0; k <= i; ++k) {
        const double cs_k = h * h * (double)(j - 1 - k);
        const double c2_k = cs_k + P.tgap * h;
        Kb[tri(i, k)] += rv * (h * h) + r1 * (cs_i * cs_k) + r2 * (c2_i * c2_k);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NH; ++i)
#pragma unroll
    for (int k = 0; k <= i; ++k) K[tri(2 * i + 1, 2 * k + 1)] += Kb[tri(i, k)];

  // Cholesky K = L L' in place (lower), keeping reciprocal diagonals
  double rdiag[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double d = K[tri(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fma(-K[tri(j, k)], K[tri(j, k)], d);
    const double rs = rsqrt(d);
    K[tri(j, j)] = d * rs;
    rdiag[j] = rs;
#pragma unroll
    for (int i = j + 1; i < NV; ++i) {
      double s = K[tri(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fma(-K[tri(i, k)], K[tri(j, k)], s);
      K[tri(i, j)] = s * rs;
    }
  }
  // Linv (lower) in place of a second packed array
  double Li[NTRI];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    Li[tri(j, j)] = rdiag[j];
#pragma unroll
    for (int i = j + 1; i < NV; ++i) {
      double s = 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) s = fma(-K[tri(i, k)], Li[tri(k, j)], s);
      Li[tri(i, j)] = s * rdiag[i];
    }
  }
  // Kinv = Linv' Linv
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = i; k < NV; ++k) s = fma(Li[tri(k, i)], Li[tri(k, j)], s);
      pb.Kinv[tri(i, j)] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// Row sweeps.  `f(row_index, v_ref, lo, hi, inrm)` is applied to every live row; the functor decides what
// to do.  Constant rows (obstacle R1 at step 1) are not rows of the QP.
// ------------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void for_each_row(const DevParams& P, Problem& pb, F&& f) {
#pragma unroll
  for (int i = 0; i < NV; ++i) f(i, pb.vb[i], P.umin[i & 1], P.umax[i & 1], 1.0);
#pragma unroll
  for (int r = 0; r < M_LANE; ++r) f(ROW_LANE + r, pb.vl[r], -P.sld - pb.lane_c[r], P.sld - pb.lane_c[r], pb.lane_inrm[r]);
#pragma unroll
  for (int j = 0; j < NH; ++j) f(ROW_V + j, pb.vv[j], -pb.x0[4], BIG, P.inrm_v[j]);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (k < pb.n_obs) {
#pragma unroll
      for (int j = 0; j < NH; ++j) {
        if (j > 0) f(ROW_OBS + 10 * k + 2 * j, pb.vo[k][j][0], -BIG, pb.base[k][j] - P.obs_safe, P.inrm_r1[j]);
        f(ROW_OBS + 10 * k + 2 * j + 1, pb.vo[k][j][1], -BIG, pb.base[k][j] - P.tgap * pb.x0[4], P.inrm_r2[j]);
      }
    }
  }
}

// out += A' w, where w is given per row type
struct RowW {
  double b[NV], l[M_LANE], v[NH], o[2][NH][2];
};

__device__ __forceinline__ void at_mul(const DevParams& P, const Problem& pb, const RowW& w, double (&out)[NV]) {
  const double h = P.h;
#pragma unroll
  for (int i = 0; i < NV; ++i) out[i] += w.b[i];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int L = 2 * (jj + 1);
    double wD = 0.0, wO = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) { wD += w.l[3 * jj + a]; wO = fma(P.alpha_lane[a], w.l[3 * jj + a], wO); }
#pragma unroll
    for (int c = 0; c < NV; ++c)
      if (c < L) out[c] = fma(pb.D[jj][c], wD, fma(pb.O[jj][c], wO, out[c]));
  }
  double Tv[NH], Ts[NH];
#pragma unroll
  for (int j = 0; j < NH; ++j) { Tv[j] = w.v[j]; Ts[j] = 0.0; }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (k < pb.n_obs) {
#pragma unroll
      for (int j = 0; j < NH; ++j) {
        Tv[j] = fma(P.tgap, w.o[k][j][1], Tv[j]);
        Ts[j] += ((j > 0) ? w.o[k][j][0] : 0.0) + w.o[k][j][1];
      }
    }
  }
  // b_i += h sum_{j>i} Tv[j] + h^2 sum_{j>=i+2} (j-1-i) Ts[j]      (j is 1-based; arrays are j-1)
  double sv = 0.0, ps = 0.0, cs = 0.0;
#pragma unroll
  for (int i = NH - 1; i >= 0; --i) {
    sv += Tv[i];             // sum_{j-1 >= i}  <=> j > i
    out[2 * i + 1] = fma(h, sv, fma(h * h, cs, out[2 * i + 1]));
    ps += Ts[i];             // P_{j=i+1} = sum_{m>=i+1} Ts[m-1]
    cs += ps;                // C_{i-1} = sum_{j>=i+1} P_j
  }
}

// zt = A x per row type
__device__ __forceinline__ void a_mul(const DevParams& P, const Problem& pb, const double (&x)[NV], RowW& zt) {
  const double h = P.h;
#pragma unroll
  for (int i = 0; i < NV; ++i) zt.b[i] = x[i];
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const int L = 2 * (jj + 1);
    double dj = 0.0, oj = 0.0;
#pragma unroll
    for (int c = 0; c < NV; ++c)
      if (c < L) { dj = fma(pb.D[jj][c], x[c], dj); oj = fma(pb.O[jj][c], x[c], oj); }
#pragma unroll
    for (int a = 0; a < 3; ++a) zt.l[3 * jj + a] = fma(P.alpha_lane[a], oj, dj);
  }
  double cum = 0.0, S = 0.0;
#pragma unroll
  for (int j = 0; j < NH; ++j) {
    // S_{j+1} = h sum_{m=1..j} (v_m - v0)
    if (j > 0) S = fma(h, zt.v[j - 1], S);
    cum = fma(h, x[2 * j + 1], cum);
    zt.v[j] = cum;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      zt.o[k][j][0] = S;
      zt.o[k][j][1] = fma(P.tgap, cum, S);
    }
  }
}

__device__ __forceinline__ void sym_mul(const double (&Kinv)[NTRI], const double (&r)[NV], double (&x)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) acc = fma(Kinv[i >= j ? tri(i, j) : tri(j, i)], r[j], acc);
    x[i] = acc;
  }
}

// row-type accessor into a RowW by global row index (compile-time after unrolling)
__device__ __forceinline__ double& roww(RowW& w, int r) {
  if (r < ROW_LANE) return w.b[r];
  if (r < ROW_V) return w.l[r - ROW_LANE];
  if (r < ROW_OBS) return w.v[r - ROW_V];
  const int q = r - ROW_OBS;
  return w.o[q / 10][(q % 10) / 2][q & 1];
}

struct SegStats { double rp, rd, nd, atdy, sup, bad; };

// One ADMM iteration.  LAST additionally evaluates residuals, certificate quantities and updates the ladder.
template <bool LAST>
__device__ __forceinline__ void admm_iter(const DevParams& P, Problem& pb, SegStats& st) {
  RowW w;
  for_each_row(P, pb, [&](int r, double& v, double lo, double hi, double inrm) {
    const double z = clipd(v, lo, hi);
    roww(w, r) = row_rho(P, pb.E, r, inrm) * (2.0 * z - v);
  });
  if (pb.n_obs < 2) {
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (k >= pb.n_obs) {
#pragma unroll
        for (int j = 0; j < NH; ++j) { w.o[k][j][0] = 0.0; w.o[k][j][1] = 0.0; }
      }
  }
  double rhs[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) rhs[i] = -pb.q[i];
  at_mul(P, pb, w, rhs);
  sym_mul(pb.Kinv, rhs, pb.x);
  RowW zt;
  a_mul(P, pb, pb.x, zt);
  if (!LAST) {
    for_each_row(P, pb, [&](int r, double& v, double lo, double hi, double inrm) {
      const double z = clipd(v, lo, hi);
      v = fma(P.relax, roww(zt, r) - z, v);
    });
  } else {
    RowW t1, t2;   // t1 = rho (z - zt) + dy ; t2 = dy
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int j = 0; j < NH; ++j) { t1.o[k][j][0] = t1.o[k][j][1] = 0.0; t2.o[k][j][0] = t2.o[k][j][1] = 0.0; }
    double rp = 0.0, nd = 0.0, sup = 0.0, bad = 0.0;
    unsigned long long act = 0ull;
    for_each_row(P, pb, [&](int r, double& v, double lo, double hi, double inrm) {
      const int e = pb.E.get(r);
      const double rho = P.lad[e] * inrm;
      const double z = clipd(v, lo, hi);
      const double ztr = roww(zt, r);
      const double vn = fma(P.relax, ztr - z, v);
      const double zn = clipd(vn, lo, hi);
      const double dy = rho * ((vn - zn) - (v - z));
      rp = fmax(rp, fabs(ztr - zn));
      nd = fmax(nd, fabs(dy));
      if (dy > 0.0) { if (hi < BIG) sup = fma(hi, dy, sup); else bad = fmax(bad, dy); }
      else if (dy < 0.0) { if (lo > -BIG) sup = fma(lo, dy, sup); else bad = fmax(bad, -dy); }
      roww(t1, r) = fma(rho, z - ztr, dy);
      roww(t2, r) = dy;
      // ladder with hysteresis
      const bool a_now = (vn < lo) || (vn > hi);
      const bool a_prev = (pb.act_prev >> r) & 1ull;
      double vnew = vn;
      if (a_now && (a_prev || !P.hysteresis) && e < P.n_rung - 1) {
        const int e2 = (e + P.up_step < P.n_rung - 1) ? e + P.up_step : P.n_rung - 1;
        pb.E.set(r, e2);
        vnew = fma(P.lad[e] / P.lad[e2], vn - zn, zn);   // keep (z, y): v' = z + (rho/rho') (v - z)
      } else if (!a_now && (!a_prev || !P.hysteresis) && e > 0) {
        pb.E.set(r, P.drop_all ? 0 : e - 1);           // inactive: v == z, nothing to rescale
      }
      if (a_now) act |= (1ull << r);
      v = vnew;
    });
    pb.act_prev = act;
    double o1[NV], o2[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { o1[i] = 0.0; o2[i] = 0.0; }
    at_mul(P, pb, t1, o1);
    at_mul(P, pb, t2, o2);
    double rd = 0.0, atdy = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) { rd = fmax(rd, fabs(o1[i])); atdy = fmax(atdy, fabs(o2[i])); }
    st.rp = rp; st.rd = rd; st.nd = nd; st.atdy = atdy; st.sup = sup; st.bad = bad;
  }
}

// ------------------------------------------------------------------------------------------------
// Whole solve for one problem.  Warp-uniform loops (votes) so that divergence only idles lanes.
// ------------------------------------------------------------------------------------------------
struct SolveOut { int status, rounds, iters; bool const_infeasible; };

__device__ __forceinline__ void init_admm_state(const DevParams& P, Problem& pb) {
  // z = clip(A U), y = 0  ->  v = z ; all rows on the initial rung
#pragma unroll
  for (int i = 0; i < 6; ++i) pb.E.w[i] = 0x11111111u * (unsigned)P.e_init;
  pb.act_prev = 0ull;
  RowW zt;
  a_mul(P, pb, pb.U, zt);
  for_each_row(P, pb, [&](int r, double& v, double lo, double hi, double inrm) { v = clipd(roww(zt, r), lo, hi); });
}

__device__ __forceinline__ SolveOut solve_one(const DevTable& T, const DevParams& P, Problem& pb, bool live) {
  // obstacle row offsets
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int j = 0; j < NH; ++j)
      pb.base[k][j] = (pb.obs[k][0] + pb.obs[k][1] * ((j + 1) * P.h)) - pb.x0[0] - (j + 1) * P.h * pb.x0[4];
  warm_start(T, P, pb.x0, pb.obs, pb.n_obs, pb.U);
#pragma unroll
  for (int i = 0; i < NV; ++i) pb.U[i] = clipd(pb.U[i], P.umin[i & 1], P.umax[i & 1]);  // scipy clips x0 to the bounds

  SolveOut out{MPCB_MAXITER, 0, 0, false};
  bool done = !live;
  bool infeasible = false;
  bool first = true;
  for (int round = 0; round < P.max_rounds; ++round) {
    if (MPCB_ALL(done)) break;
    if (!done) {
      double cviol;
      linearise(T, P, pb, cviol);
      if (first) { init_admm_state(P, pb); first = false; }
      if (cviol > P.feas_tol) { infeasible = true; out.const_infeasible = true; }
      out.rounds++;
    }
    bool conv = done;
    bool cert = false;
    for (int seg = 0; seg < P.max_segments; ++seg) {
      if (MPCB_ALL(conv)) break;
      if (!conv) {
        factor(P, pb);
        SegStats st;
        for (int it = 0; it < P.segment_iters - 1; ++it) admm_iter<false>(P, pb, st);
        admm_iter<true>(P, pb, st);
        out.iters += P.segment_iters;
#ifdef MPCB_TRACE
        if (getenv("MPCB_TRACE")) {
          printf("  r%d s%d rp %.2e rd %.2e nd %.2e act %012llx E", round, seg, st.rp, st.rd, st.nd, pb.act_prev);
          for (int r = 0; r < M_ROWS; ++r) printf("%d", pb.E.get(r));
          printf(" x");
          for (int i = 0; i < NV; ++i) printf(" %.4f", pb.x[i]);
          printf("\n");
        }
#endif
        if (st.rp <= P.eps_p && st.rd <= P.eps_d) conv = true;
        else if (P.trust_cert && st.nd > 1e-9 && st.atdy <= P.eps_inf * st.nd && st.sup < -P.eps_inf * st.nd &&
                 st.bad <= P.eps_inf * st.nd) { conv = true; cert = true; }
      }
    }
    if (!done) {
      double step = 0.0;
#pragma unroll
      for (int i = 0; i < NV; ++i) { step = fmax(step, fabs(pb.x[i] - pb.U[i])); pb.U[i] = pb.x[i]; }
      if (cert) { infeasible = true; done = true; }
      else if (conv && step < P.step_tol) { done = true; out.status = 0; }
      else if (!conv && !P.trust_cert) done = true;   // first pass: a QP it cannot close goes to the robust pass
    }
  }
  if (infeasible) out.status = 2;
  // the returned controls always respect the box (a no-op for converged problems; non-converged or infeasible
  // ADMM iterates may sit slightly outside).  SLSQP treats the bounds as hard in the same way.
#pragma unroll
  for (int i = 0; i < NV; ++i) pb.U[i] = clipd(pb.U[i], P.umin[i & 1], P.umax[i & 1]);
  return out;
}

}  // namespace mpcb
