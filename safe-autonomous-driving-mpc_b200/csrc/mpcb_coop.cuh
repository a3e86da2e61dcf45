// mpcb_coop.cuh -- the same SQP / ADMM algorithm as mpcb_solver.cuh, executed by ONE WARP PER PROBLEM.
//
// Why a second execution shape: one thread per problem maximises throughput on a full batch but a single thread walks
// ~1,000 instructions per ADMM iteration, so the few hard problems of the robust pass (and a B = 1 call) take as long
// as their serial instruction chain.  Here a warp runs alone on its scheduler and every dependent instruction costs
// its full latency, so the code is arranged for short dependent chains:
//   * the 41 rows of the QP are spread over the lanes (lane l owns rows l and l + 32 as dense 10-vectors, with their
//     ADMM state in registers); without a second obstacle nobody has a second row and that work is skipped;
//   * A'w: every lane publishes the weights of its rows, lane (i, part) adds up column i over a third of the rows,
//     two shuffles combine the thirds; every lane then evaluates its row values as g_r . (A'w - q), g_r = K^-1 a_r
//     precomputed per factorisation: one exchange and one broadcast per iteration; x is read off the box rows;
//   * factorisation: K accumulated entry-per-lane in straight-line row blocks, Cholesky redundantly in registers;
//   * linearisation by the whole warp (coop_linearise): sensitivity columns and H entries over the lanes, the six
//     independent table lookups of a rollout on six lanes at once (likewise the five of the warm start);
//   * residual / certificate tests as warp votes, so the lanes take the same branches; warps are independent (no CTA
//     barrier) and pull their problems from a device counter (mpcb_api.cu).
// The screen, the warm-start logic and the final evaluation reuse the thread-level code on lane 0 with a
// thread-private store placed in shared memory.
#pragma once
#include "mpcb_solver.cuh"

namespace mpcb {

#ifdef MPCB_COOP_PROFILE
// development: cycles per phase, summed over warps, and the phase split of the slowest warp
__device__ unsigned long long g_coop_sum[8];
__device__ unsigned long long g_coop_max[8];
#define CPROF_T(var) const long long var = clock64()
#define CPROF_ADD(slot, t0, t1) prof[slot] += (t1) - (t0)
#else
#define CPROF_T(var)
#define CPROF_ADD(slot, t0, t1)
#endif

constexpr int CW_K = 2;                 // rows per lane
constexpr int CW_ROWS = 32 * CW_K;      // padded row count (rows >= M_ROWS are null)

struct WarpShared {
  double buf[Store<1, 0u>::LOCAL];       // lane 0's thread-level store (H, q, D, O, lane_c, ...)
  Problem pb;
  double A[CW_ROWS][NV + 1];            // dense rows (+1 pad: conflict-free column reads)
  double rho[CW_ROWS];
  double K[NV][NV + 1];                 // K, then its Cholesky factor (lower)
  double rdiag[NV];
  alignas(16) double tv[2][CW_ROWS];    // row weights published for the A't products
  double vec[2][NV];                    // reduced vector / iterate, broadcast to the lanes
  double jv[2][3][NV];                  // linearisation: residual Jacobian rows of a step (d, o, v), double buffered
  double xh[NH + 1][2];                 // linearisation: (d_j, o_j) of the nominal rollout
  double lk[NH + 1][8];                 // linearisation: reference values and slopes at s_0 .. s_5
  double ur[NH][2];                     // warm start: reference controls at the five probe positions
  double hot[NV];                       // hot start of the first pass (mpcb_api.cu), when the caller supplies one
  double scal[4];                       // lane-0 scalars broadcast through shared memory
  int iflag[2];
};

// Cholesky factor held in registers by every lane (the 10x10 factorisation is cheaper done redundantly by all lanes
// than passed around: no barrier, no shared-memory round trip per column)
struct RegFactor {
  double L[NTRI];
  double rdiag[NV];
};

__device__ __forceinline__ void chol_regs(RegFactor& F) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double d = F.L[tri(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fma(-F.L[tri(j, k)], F.L[tri(j, k)], d);
    const double rs = rsqrt(d);
    F.L[tri(j, j)] = d * rs;
    F.rdiag[j] = rs;
#pragma unroll
    for (int i = j + 1; i < NV; ++i) {
      double t = F.L[tri(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) t = fma(-F.L[tri(i, k)], F.L[tri(j, k)], t);
      F.L[tri(i, j)] = t * rs;
    }
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// max over the warp of a non-negative finite double: its bit pattern orders like an unsigned 64-bit integer, so two
// 32-bit REDUX instructions replace five shuffle / compare steps
__device__ __forceinline__ double warp_max_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}

// dense coefficients, bounds and 1/|a|^2 of row r at the current linearisation (same rows as for_rows)
__device__ __forceinline__ void coop_row(const DevParams& P, const Store<1, 0u>& st, const Problem& pb, int r,
                                         double (&a)[NV], double& lo, double& hi, double& inrm, bool& exists) {
  const double h = P.h;
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = 0.0;
  lo = -BIG; hi = BIG; inrm = 0.0; exists = false;
  if (r < ROW_LANE) {
#pragma unroll
    for (int i = 0; i < NV; ++i) if (i == r) a[i] = 1.0;
    if (r & 1) { lo = st.blo[r >> 1]; hi = st.bhi[r >> 1]; } else { lo = P.umin[0]; hi = P.umax[0]; }
    inrm = 1.0; exists = true;
  } else if (r < ROW_V) {
    const int q = r - ROW_LANE, jj = q >> 1;
    const double al = (q & 1) ? P.alpha_lane[2] : 0.0;
#pragma unroll
    for (int c = 0; c < NV - 2; ++c)
      if (c < 2 * (jj + 1)) a[c] = fma(al, st.O[doff(jj) + c], st.D[doff(jj) + c]);
    lo = -P.sld - st.lane_c[q]; hi = P.sld - st.lane_c[q]; inrm = st.lane_inrm[q]; exists = true;
  } else if (r < ROW_OBS) {
    const int j = r - ROW_V + 1;
#pragma unroll
    for (int i = 0; i < NH; ++i) if (i < j) a[2 * i + 1] = h;
    lo = st.lov[j - 1]; inrm = P.inrm_v[j - 1]; exists = true;
  } else if (r < M_ROWS) {
    const int q = (r - ROW_OBS) % N_OBSROW, k = (r - ROW_OBS) / N_OBSROW;
    const bool r2 = q >= 4;
    const int j = r2 ? q - 3 : q + 2;
#pragma unroll
    for (int i = 0; i < NH; ++i) {
      double c = (i < j - 1) ? h * h * (double)(j - 1 - i) : 0.0;
      if (r2 && i < j) c += P.tgap * h;
      a[2 * i + 1] = c;
    }
    hi = st.hio[N_OBSROW * k + q];
    inrm = r2 ? P.inrm_r2[j - 1] : P.inrm_r1[j - 1];
    exists = k < pb.n_obs;
  }
}


// ------------------------------------------------------------------------------------------------
// Linearisation by the whole warp: same arithmetic as linearise() of mpcb_solver.cuh, entry by entry in the same
// order, spread over the lanes.  The nominal rollout and the table lookups run redundantly on every lane (same
// instruction stream, broadcast loads); lane c < 10 carries column c of the sensitivities d(d_j)/dU, d(o_j)/dU and
// component c of the gradient; the 55 entries of H belong to the lanes like the entries of K (ki, kj); the residual
// Jacobian rows of a step reach their consumers through shared memory.  Fills st.H, q, D, O, lane_c, lane_inrm.
// ------------------------------------------------------------------------------------------------
template <class ST>
__device__ void coop_linearise(const DevTable& T, const DevParams& P, WarpShared& ws, const ST& st, int lane,
                               const int (&ki)[2], const int (&kj)[2], double& const_viol) {
  const double h = P.h;
  Problem& pb = ws.pb;
  const int c = lane, ci = lane >> 1;
  const bool cb = (lane & 1) != 0, col = lane < NV;
  double Hm[2];
#pragma unroll
  for (int m = 0; m < 2; ++m) Hm[m] = (lane + 32 * m < NTRI && ki[m] == kj[m]) ? 2.0 * P.wu[ki[m] & 1] : 0.0;
  double g = col ? 2.0 * P.wu[c & 1] * pb.U[c] : 0.0;
  double X[5] = {pb.x0[0], pb.x0[1], pb.x0[2], pb.x0[3], pb.x0[4]};
  double dD = 0.0, dO = 0.0;
  double val[4] = {0.0, 0.0, 0.0, 0.0}, slope[4] = {0.0, 0.0, 0.0, 0.0};
  double cv = 0.0;
  // s_j and v_j do not depend on the reference values (s' = v, v' = b): the six table lookups of the rollout are
  // independent of one another, so lanes 0..5 do one each -- one memory round trip instead of six in a row
  {
    double sj = X[0], vj = X[4], s_mine = X[0];
#pragma unroll
    for (int j = 0; j <= NH; ++j) {
      if (lane == j) s_mine = sj;
      if (j < NH) { sj = sj + h * vj; vj = vj + h * pb.U[2 * j + 1]; }
    }
    if (lane <= NH) {
      int hint = pb.hint[lane];
      lookup_state_hint(T, s_mine, val, slope, hint);
      pb.hint[lane] = hint;
#pragma unroll
      for (int q = 0; q < 4; ++q) { ws.lk[lane][q] = val[q]; ws.lk[lane][4 + q] = slope[q]; }
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j <= NH; ++j) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { val[q] = ws.lk[j][q]; slope[q] = ws.lk[j][4 + q]; }
    if (lane == 0) { ws.xh[j][0] = X[1]; ws.xh[j][1] = X[2]; }
    if (j >= 1) {
      // residual rows of step j (accumulate_step<j>): this lane's column of J_d, J_o, J_v
      const double ds = (ci < j - 1) ? h * h * (double)(j - 1 - ci) : 0.0;
      double Jd = 0.0, Jo = 0.0, Jv = 0.0;
      if (col && cb && ci < j) Jv = h - slope[3] * ds;
      if (col && ci < j - 1) {
        Jd = cb ? dD - slope[0] * ds : dD;
        Jo = cb ? dO - slope[1] * ds : dO;
      }
      const double rd = X[1] - val[0], ro = X[2] - val[1], rv = X[4] - val[3];
      if (col && ci < j - 1) g += 2.0 * (P.wd * rd * Jd + P.wo * ro * Jo);
      if (col && cb && ci < j) g += 2.0 * P.wv * rv * Jv;
      double (*jv)[NV] = ws.jv[j & 1];
      if (col) { jv[0][c] = Jd; jv[1][c] = Jo; jv[2][c] = Jv; }
      // lane rows of step j (lane_rows<j>): sensitivities as they stand before the next advance
      if (j >= 2 && col && c < 2 * (j - 1)) { st.D[doff(j - 2) + c] = dD; st.O[doff(j - 2) + c] = dO; }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        if (lane + 32 * m < NTRI) {
          const int p = ki[m], q = kj[m];
          if (j >= 2) {
            Hm[m] = fma((2.0 * P.wd) * jv[0][p], jv[0][q], Hm[m]);
            Hm[m] = fma((2.0 * P.wo) * jv[1][p], jv[1][q], Hm[m]);
          }
          Hm[m] = fma((2.0 * P.wv) * jv[2][p], jv[2][q], Hm[m]);
        }
      }
      if (j == 1) {
        cv = dmax(cv, fabs(X[1]) - P.sld);
        cv = dmax(cv, fabs(X[1] + P.alpha_lane[2] * X[2]) - P.sld);
      }
    }
    if (j < NH) {
      // advance j -> j + 1
      const double d = X[1], o = X[2], k = X[3], v = X[4];
      const double kk = k - val[2];
      if (col && ci < j) {
        const double ds = (ci < j - 1) ? h * h * (double)(j - 1 - ci) : 0.0;
        double nD, nO;
        if (!cb) {
          nD = dD + h * (v * dO);
          nO = dO + h * (v * h);
        } else {
          nD = dD + h * (h * o + v * dO);
          nO = dO + h * (h * kk - v * (slope[2] * ds));
        }
        dD = nD;
        dO = nO;
      }
      X[0] = X[0] + h * v;
      X[1] = d + h * (v * o);
      X[2] = o + h * (v * kk);
      X[3] = k + h * pb.U[2 * j];
      X[4] = v + h * pb.U[2 * j + 1];
    }
  }
#pragma unroll
  for (int m = 0; m < 2; ++m)
    if (lane + 32 * m < NTRI) st.H[lane + 32 * m] = Hm[m];
  __syncwarp();
  // lane rows: offsets and norms (lane r < 8 takes row r), then q = g - H U on lanes 0..9
  if (lane < N_LANE) {
    const int jj = lane >> 1, L = 2 * (jj + 1), o = doff(jj);
    double dU = 0.0, oU = 0.0, dd = 0.0, dox = 0.0, oo = 0.0;
#pragma unroll
    for (int cc = 0; cc < NV - 2; ++cc) {
      if (cc < L) {
        const double Dc = st.D[o + cc], Oc = st.O[o + cc], Uc = pb.U[cc];
        dU = fma(Dc, Uc, dU);
        oU = fma(Oc, Uc, oU);
        dd = fma(Dc, Dc, dd);
        dox = fma(Dc, Oc, dox);
        oo = fma(Oc, Oc, oo);
      }
    }
    const double al = (lane & 1) ? P.alpha_lane[2] : 0.0;
    st.lane_c[lane] = (ws.xh[jj + 2][0] + al * ws.xh[jj + 2][1]) - (dU + al * oU);
    const double n2 = dd + 2.0 * al * dox + al * al * oo;
    st.lane_inrm[lane] = 1.0 / dmax(n2, NRM2_FLOOR);
  }
  if (col) {
    double acc = g;
#pragma unroll
    for (int j = 0; j < NV; ++j) acc = fma(-st.H[c >= j ? tri(c, j) : tri(j, c)], pb.U[j], acc);
    st.q[c] = acc;
  }
  const_viol = cv;
  __syncwarp();
}

template <bool FIRST_PASS>
__device__ SolveOut coop_solve(const DevTable& T, const DevParams& P, WarpShared& ws, int lane, bool hot = false) {
  const unsigned FULL = 0xffffffffu;
  const Store<1, 0u> st(nullptr, ws.buf);
  Problem& pb = ws.pb;
  const Policy& pl = P.pol[FIRST_PASS ? 1 : 0];
  SolveOut out{MPCB_MAXITER, 0, 0, false};
#ifdef MPCB_COOP_PROFILE
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // 0 prologue, 1 linearise, 2 rows, 3 factor, 4 iterations, 5 round end
#endif
  CPROF_T(tp0);
  // warm start: the five reference-control lookups (binary searches in the table) at once on five lanes
  if (lane < NH) {
    double s_cur = pb.x0[0], s_mine = pb.x0[0];
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      if (lane == j) s_mine = s_cur;
      s_cur += pb.x0[4] * P.h;
    }
    int hint = 0;
    double u2[2];
    lookup_control(T, s_mine, u2, hint);
    ws.ur[lane][0] = u2[0];
    ws.ur[lane][1] = u2[1];
    pb.hint[lane] = hint;
  }
  __syncwarp();
  if (lane == 0) {
    ws.iflag[0] = prologue(T, P, pb, st, ws.ur) ? 1 : 0;
    if (FIRST_PASS && hot) {               // start from the supplied plan, clipped like any start (solve_one does the same)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const double u = clipd(ws.hot[i], P.umin[i & 1], P.umax[i & 1]);
        pb.U[i] = (i & 1) ? clipd(u, st.blo[i >> 1], st.bhi[i >> 1]) : u;
      }
    }
  }
  __syncwarp();
  CPROF_T(tp1);
  CPROF_ADD(0, tp0, tp1);
  const bool screened = ws.iflag[0] != 0;
  bool infeasible = screened;
  out.const_infeasible = screened;
  bool done = false, first = true;
  double step_prev = 1e30;
  int fails = 0;
  const int max_rounds = FIRST_PASS ? P.fast_max_rounds : P.max_rounds;

  // per-lane row state
  double a[CW_K][NV], v[CW_K], rho[CW_K], lo[CW_K], hi[CW_K], inrm[CW_K];
  int e[CW_K];
  bool ex[CW_K], aprev[CW_K];
  double x[NV];       // replicated iterate (refreshed at the end of every segment)
  double g[CW_K][NV]; // K^-1 a_r for this lane's rows: (A x)_r = g_r . (A'w - q)
#pragma unroll
  for (int i = 0; i < NV; ++i) x[i] = 0.0;
#pragma unroll
  for (int k = 0; k < CW_K; ++k) { v[k] = 0.0; rho[k] = 0.0; e[k] = 0; aprev[k] = false; }

  // o[i] = sum over all rows of a_r[i] t_r for one or two row-weight vectors at once, result on lanes 0..9 (lane i
  // gets component i).  Every lane publishes the weights of its rows; lane (i, part), i = lane % 10, part = lane / 10
  // < 3, adds up column i of A' over a third of the rows (straight-line, loads issued ahead of the FMAs), two
  // shuffles combine the thirds.  (Rows 41 and up are null: zero coefficients, zero weights.)
  auto at_reduce = [&](const double (&t0)[CW_K], const double (&t1)[CW_K], bool two, double& o0, double& o1) {
    ws.tv[0][lane] = t0[0];
    ws.tv[0][lane + 32] = t0[1];
    if (two) { ws.tv[1][lane] = t1[0]; ws.tv[1][lane + 32] = t1[1]; }
    __syncwarp();
    const int i = lane % NV, part = lane / NV;
    double s0 = 0.0, s0b = 0.0, s1 = 0.0, s1b = 0.0;
    if (part < 3) {
      const int r0 = part * 14;
#pragma unroll
      for (int m = 0; m < 14; m += 2) {
        const double2 ta = *reinterpret_cast<const double2*>(&ws.tv[0][r0 + m]);
        const double c0 = ws.A[r0 + m][i], c1 = ws.A[r0 + m + 1][i];
        s0 = fma(c0, ta.x, s0);
        s0b = fma(c1, ta.y, s0b);
        if (two) {
          const double2 tb = *reinterpret_cast<const double2*>(&ws.tv[1][r0 + m]);
          s1 = fma(c0, tb.x, s1);
          s1b = fma(c1, tb.y, s1b);
        }
      }
      s0 += s0b;
      s1 += s1b;
    }
    o0 = s0 + __shfl_down_sync(FULL, s0, 10) + __shfl_down_sync(FULL, s0, 20);
    o1 = two ? s1 + __shfl_down_sync(FULL, s1, 10) + __shfl_down_sync(FULL, s1, 20) : 0.0;
    __syncwarp();
  };

  // packed-triangle entries of K this lane accumulates (two passes over the 55 entries)
  int ki[2], kj[2];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int idx = lane + 32 * m;
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= idx) ++i;
    ki[m] = i;
    kj[m] = idx - i * (i + 1) / 2;
  }

  for (int round = 0; round < max_rounds && !done; ++round) {
    CPROF_T(tl0);
    double cviol;
    coop_linearise(T, P, ws, st, lane, ki, kj, cviol);
    CPROF_T(tl1);
    CPROF_ADD(1, tl0, tl1);
    if (cviol > P.feas_tol) { infeasible = true; out.const_infeasible = true; }
    out.rounds++;
    // dense rows of this round
#pragma unroll
    for (int k = 0; k < CW_K; ++k) {
      const int r = lane + 32 * k;
      coop_row(P, st, pb, r, a[k], lo[k], hi[k], inrm[k], ex[k]);
#pragma unroll
      for (int i = 0; i < NV; ++i) ws.A[r][i] = a[k][i];
    }
    if (first) {
      // cold ADMM state at U: z = clip(A U), y = 0 -> v = z ; all rows on the initial rung
#pragma unroll
      for (int i = 0; i < NV; ++i) x[i] = pb.U[i];
#pragma unroll
      for (int k = 0; k < CW_K; ++k) {
        double zt = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) zt = fma(a[k][i], x[i], zt);
        v[k] = clipd(zt, lo[k], hi[k]);
        e[k] = pl.e_init;
        aprev[k] = false;
        if (FIRST_PASS && hot && ex[k] && (zt <= lo[k] + P.feas_tol || zt >= hi[k] - P.feas_tol)) {
          e[k] = pl.n_rung - 1;             // hot start: a row on its bound starts as active (two-level policy: rung "on")
          aprev[k] = true;
        }
      }
      first = false;
    }
    CPROF_T(tr1);
    CPROF_ADD(2, tl1, tr1);
    const double loosen = (FIRST_PASS || P.qp_forcing <= 0.0) ? 1.0
                          : dmax(1.0, (step_prev * P.qp_forcing < P.qp_eps_loose ? step_prev * P.qp_forcing : P.qp_eps_loose) / P.eps_p);
    const double eps_p = P.eps_p * loosen, eps_d = P.eps_d * loosen;
    bool conv = false, cert = false;
    for (int seg = 0; seg < pl.max_segments && !conv; ++seg) {
      // ---- factor: K = H + A' diag(rho) A, Cholesky, K^-1 rows ------------------------------------------------
      CPROF_T(tf0);
#pragma unroll
      for (int k = 0; k < CW_K; ++k) {
        rho[k] = ex[k] ? pl.lad[e[k]] * inrm[k] : 0.0;
        ws.rho[lane + 32 * k] = rho[k];
      }
      __syncwarp();
      // K = H + sum_r rho_r a_r a_r': box rows only touch the diagonal; the other rows in straight-line blocks (lane
      // rows, speed rows, one block per present obstacle) so that a block's loads are issued ahead of its FMAs
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int idx = lane + 32 * m;
        if (idx < NTRI) {
          const int i = ki[m], j = kj[m];
          double acc = st.H[idx] + (i == j ? ws.rho[i] : 0.0), acc2 = 0.0;
          auto kblock = [&](auto r0t, auto r1t) {
            constexpr int R0 = decltype(r0t)::value, R1 = decltype(r1t)::value;
#pragma unroll
            for (int r = R0; r < R1; r += 2) {
              acc = fma(ws.rho[r] * ws.A[r][i], ws.A[r][j], acc);
              if (r + 1 < R1) acc2 = fma(ws.rho[r + 1] * ws.A[r + 1][i], ws.A[r + 1][j], acc2);
            }
          };
          kblock(std::integral_constant<int, ROW_LANE>{}, std::integral_constant<int, ROW_OBS>{});
          if (pb.n_obs >= 1) kblock(std::integral_constant<int, ROW_OBS>{}, std::integral_constant<int, ROW_OBS + N_OBSROW>{});
          if (pb.n_obs == 2) kblock(std::integral_constant<int, ROW_OBS + N_OBSROW>{}, std::integral_constant<int, M_ROWS>{});
          acc += acc2;
          ws.K[i][j] = acc;
          ws.K[j][i] = acc;
        }
      }
      __syncwarp();
      RegFactor F;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) F.L[tri(i, j)] = ws.K[i][j];
      chol_regs(F);
      const bool two = pb.n_obs == 2;     // rows 32..40 are exactly the second obstacle's: no second rows without it
#pragma unroll
      for (int k = 0; k < CW_K; ++k) {
        if (k == 0 || two) {
          double t[NV];
#pragma unroll
          for (int i = 0; i < NV; ++i) t[i] = a[k][i];
          chol_solve(F, t, g[k]);
        } else {
#pragma unroll
          for (int i = 0; i < NV; ++i) g[k][i] = 0.0;
        }
      }
      double zt0_last = 0.0;
      // ---- iterations ---------------------------------------------------------------------------------------------
      CPROF_T(tf1);
      CPROF_ADD(3, tf0, tf1);
      bool qp_ok = false, cert_ok = false;
      for (int it = 0; it < pl.segment_iters; ++it) {
        const bool check = (it == pl.segment_iters - 1);
        double z[CW_K], w[CW_K];
#pragma unroll
        for (int k = 0; k < CW_K; ++k) {
          z[k] = clipd(v[k], lo[k], hi[k]);
          w[k] = rho[k] * fma(2.0, z[k], -v[k]);
        }
        double r0, r1;
        at_reduce(w, w, false, r0, r1);
        if (lane < NV) ws.vec[0][lane] = r0 - st.q[lane];
        __syncwarp();
        double rhs[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) rhs[i] = ws.vec[0][i];
        double zt[CW_K];
#pragma unroll
        for (int k = 0; k < CW_K; ++k) {
          double acc = 0.0, accb = 0.0;
          if (k == 0 || two) {
#pragma unroll
            for (int i = 0; i < NV; i += 2) { acc = fma(g[k][i], rhs[i], acc); accb = fma(g[k][i + 1], rhs[i + 1], accb); }
          }
          zt[k] = acc + accb;
        }
        zt0_last = zt[0];
        if (!check) {
#pragma unroll
          for (int k = 0; k < CW_K; ++k) v[k] = fma(pl.relax, zt[k] - z[k], v[k]);
          continue;
        }
        double t1[CW_K], t2[CW_K];
        double rp = 0.0, nd = 0.0, sup = 0.0, bad = 0.0;
#pragma unroll
        for (int k = 0; k < CW_K; ++k) {
          const double vn = fma(pl.relax, zt[k] - z[k], v[k]);
          const double zn = clipd(vn, lo[k], hi[k]);
          rp = dmax(rp, fabs(zt[k] - zn));
          t1[k] = rho[k] * (fma(2.0 - pl.relax, z[k], (pl.relax - 1.0) * zt[k]) - zn);
          const double dy = rho[k] * ((vn - zn) - (v[k] - z[k]));
          t2[k] = dy;
          if (!FIRST_PASS) {
            nd = dmax(nd, fabs(dy));
            if (dy > 0.0) { if (hi[k] < BIG) sup = fma(hi[k], dy, sup); else bad = dmax(bad, dy); }
            else if (dy < 0.0) { if (lo[k] > -BIG) sup = fma(lo[k], dy, sup); else bad = dmax(bad, -dy); }
          }
          const bool a_now = (vn < lo[k]) || (vn > hi[k]);
          double vnew = vn;
          if (a_now && (aprev[k] || !pl.hysteresis) && e[k] < pl.n_rung - 1) {
            e[k] += 1;
            vnew = fma(pl.lad_ratio[e[k]], vn - zn, zn);
          } else if (!a_now && (!aprev[k] || !pl.hysteresis) && e[k] > 0) {
            e[k] = pl.drop_all ? 0 : e[k] - 1;
          }
          aprev[k] = a_now;
          v[k] = vnew;
        }
        double o1, o2;
        at_reduce(t1, t2, !FIRST_PASS, o1, o2);
        // convergence and certificate tests as warp votes (same decisions as comparing the warp-wide maxima)
        qp_ok = __all_sync(FULL, rp <= eps_p) && __all_sync(FULL, lane < NV ? fabs(o1) <= eps_d : true);
        if (!FIRST_PASS && !qp_ok) {
          const double ndw = warp_max_nonneg(nd);
          const double thr = P.eps_inf * ndw;
          const bool c_at = __all_sync(FULL, lane < NV ? fabs(o2) <= thr : true);
          const bool c_bad = __all_sync(FULL, bad <= thr);
          const double supw = warp_sum(sup);
          cert_ok = ndw > 1e-9 && c_at && c_bad && supw < -thr;
        }
      }
      // lanes 0..9 own the box rows, whose row value is x_i itself
#pragma unroll
      for (int i = 0; i < NV; ++i) x[i] = __shfl_sync(FULL, zt0_last, i);
      out.iters += pl.segment_iters;
      CPROF_T(ti1);
      CPROF_ADD(4, tf1, ti1);
      if (qp_ok) conv = true;
      else if (cert_ok) { conv = true; cert = true; }
    }
    double step = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) step = dmax(step, fabs(x[i] - pb.U[i]));
    __syncwarp();
    if (lane < NV) { pb.U[lane] = x[lane]; pb.x[lane] = x[lane]; }
    __syncwarp();
    step_prev = step;
    if (cert) { infeasible = true; done = true; }
    else if (conv && step < P.step_tol) { done = true; out.status = 0; }
    else if (!conv && FIRST_PASS) done = true;
    else if (!conv && ++fails >= P.max_fail_rounds) done = true;
  }
  // an infeasibility verdict is final only on a point the pass actually converged to (or, in the robust pass, gave
  // up on): a first pass that could not close its QP hands the problem over whatever the screens said
  if (infeasible && (!FIRST_PASS || out.status == 0)) out.status = 2;
#ifdef MPCB_COOP_PROFILE
  if (lane == 0) {
    long long tot = 0;
    for (int i = 0; i < 6; ++i) { tot += prof[i]; atomicAdd(&g_coop_sum[i], (unsigned long long)prof[i]); }
    atomicAdd(&g_coop_sum[6], (unsigned long long)tot);
    atomicAdd(&g_coop_sum[7], 1ull);
    if ((unsigned long long)tot > atomicMax(&g_coop_max[6], (unsigned long long)tot)) {   // racy snapshot: development aid only
      for (int i = 0; i < 6; ++i) g_coop_max[i] = (unsigned long long)prof[i];
      g_coop_max[7] = (unsigned long long)out.iters * 1000ull + (unsigned long long)out.rounds;
    }
  }
#endif
  __syncwarp();
  if (lane < NV) pb.U[lane] = clipd(pb.U[lane], P.umin[lane & 1], P.umax[lane & 1]);
  __syncwarp();
  return out;
}

}  // namespace mpcb
