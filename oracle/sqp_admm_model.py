"""Batched numpy model of the GPU algorithm (Gauss-Newton SQP -> OSQP-style ADMM -> KKT polish).

TEST INFRASTRUCTURE ONLY (lives under oracle/): a vectorised CPU statement of what the CUDA kernel
computes, used by tests to localise differences (assembly values, Jacobians, QP data) and during
development to choose solver constants.  It is NOT the parity oracle (that is tracker_port +
SLSQP) and is never imported by the product package.

Reference formulation: trajectory_tracking.py:87-211 (predict / cost / constraints), restated
with exact derivatives in SURVEY.md Appendix A.
"""
from __future__ import annotations

import numpy as np

from . import tracker_port as P

H = P.DT
NV = 10
INF = 1e30


# ----------------------------------------------------------------------------------------------
def lookup(tab: P.RefTable, s):
    """Vectorised get_state(s)[1:5] and the slope of each column on the bracketing segment.
    Returns val (..,4) [d,o,k,v], slope (..,4)."""
    s = np.asarray(s, dtype=np.float64)
    i = np.clip(np.searchsorted(tab.s, s, side="left"), 1, tab.K - 1)
    lo = i - 1
    x_lo, x_hi = tab.s[lo], tab.s[i]
    wl = (s - x_lo) / (x_hi - x_lo)
    wr = (x_hi - s) / (x_hi - x_lo)
    Y = tab.X[:, 1:5]
    val = wl[..., None] * Y[i] + wr[..., None] * Y[lo]
    slope = (Y[i] - Y[lo]) / (x_hi - x_lo)[..., None]
    past = s >= tab.s_max
    val = np.where(past[..., None], Y[-1], val)
    slope = np.where(past[..., None], 0.0, slope)
    return val, slope


def lookup_control(tab: P.RefTable, s):
    s = np.asarray(s, dtype=np.float64)
    su = tab.s[: tab.Ku]
    i = np.clip(np.searchsorted(su, s, side="left"), 1, tab.Ku - 1)
    lo = i - 1
    x_lo, x_hi = su[lo], su[i]
    wl = (s - x_lo) / (x_hi - x_lo)
    wr = (x_hi - s) / (x_hi - x_lo)
    Y = tab.U[: tab.Ku]
    val = wl[..., None] * Y[i] + wr[..., None] * Y[lo]
    return np.where((s >= tab.s_max)[..., None], 0.0, val)


def warm_start(tab, x0, obs_sv, n_obs):
    """trajectory_tracking.py:223-246 batched, then clipped to the bounds as scipy does
    (_slsqp_py.py:322)."""
    B = x0.shape[0]
    s = x0[:, 0].copy()
    v = x0[:, 4]
    brake = np.zeros(B, dtype=bool)
    U = np.zeros((B, 5, 2))
    for j in range(5):
        for k in range(2):
            brake |= (n_obs > k) & ((obs_sv[:, k, 0] - s) < P.BRAKE_LOOKAHEAD)
        uref = lookup_control(tab, s)
        U[:, j, 0] = uref[:, 0]
        U[:, j, 1] = np.where(brake, P.BRAKE_GUESS, uref[:, 1])
        s = s + v * H
    U = U.reshape(B, 10)
    lb = np.tile(P.U_MIN, 5)
    ub = np.tile(P.U_MAX, 5)
    return np.clip(U, lb, ub)


def assemble(tab, x0, U):
    """Rollout + forward sensitivities.  Returns dict with X (B,6,5), dX (B,6,5,10), residuals
    r (B,15) ordered (j, [d,o,v]), Jr (B,15,10), cost (B,)."""
    B = x0.shape[0]
    X = np.zeros((B, 6, 5))
    dX = np.zeros((B, 6, 5, NV))
    X[:, 0] = x0
    Um = U.reshape(B, 5, 2)
    for j in range(5):
        s, d, o, k, v = (X[:, j, c] for c in range(5))
        val, slope = lookup(tab, s)
        kap, dkap = val[:, 2], slope[:, 2]
        ds, dd, do, dk, dv = (dX[:, j, c] for c in range(5))
        X[:, j + 1, 0] = s + H * v
        X[:, j + 1, 1] = d + H * (v * o)
        X[:, j + 1, 2] = o + H * (v * (k - kap))
        X[:, j + 1, 3] = k + H * Um[:, j, 0]
        X[:, j + 1, 4] = v + H * Um[:, j, 1]
        dX[:, j + 1, 0] = ds + H * dv
        dX[:, j + 1, 1] = dd + H * (dv * o[:, None] + v[:, None] * do)
        dX[:, j + 1, 2] = do + H * (dv * (k - kap)[:, None] + v[:, None] * (dk - dkap[:, None] * ds))
        dX[:, j + 1, 3] = dk
        dX[:, j + 1, 3, 2 * j] += H
        dX[:, j + 1, 4] = dv
        dX[:, j + 1, 4, 2 * j + 1] += H
    r = np.zeros((B, 15))
    Jr = np.zeros((B, 15, NV))
    w = np.array([P.W_D, P.W_O, P.W_V])
    cost = np.zeros(B)
    for j in range(1, 6):
        val, slope = lookup(tab, X[:, j, 0])
        for c, (col, tcol) in enumerate(((1, 0), (2, 1), (4, 3))):
            r[:, 3 * (j - 1) + c] = X[:, j, col] - val[:, tcol]
            Jr[:, 3 * (j - 1) + c] = dX[:, j, col] - slope[:, tcol][:, None] * dX[:, j, 0]
            cost += w[c] * r[:, 3 * (j - 1) + c] ** 2
    cost += 0.5 * np.sum(U ** 2, axis=1)
    return dict(X=X, dX=dX, r=r, Jr=Jr, cost=cost)


W15 = np.tile(np.array([P.W_D, P.W_O, P.W_V]), 5)
SLD = P.LANE_WIDTH / 2.0 - P.VEHICLE_RADIUS - P.SAFE_LANE_MARGIN
ALPHAS = (0.0, P.WHEELBASE / 2.0, P.WHEELBASE)

# constant sensitivity rows of v_j and s_j (j=1..5) w.r.t. U (exactly affine)
SV = np.zeros((5, NV))
SS = np.zeros((5, NV))
for _j in range(1, 6):
    for _i in range(_j):
        SV[_j - 1, 2 * _i + 1] = H
        if _i < _j - 1:
            SS[_j - 1, 2 * _i + 1] = H * H * (_j - 1 - _i)

# QP row layout (M = 47): 0-9 box, 10-21 lane (j=2..5 x 3 alphas), 22-26 v_j>=0, 27-36 obstacle 0
# (j=1..5 x [R1: gap-5, R2: gap-1.5v]), 37-46 obstacle 1.
M = 47
ROW_BOX, ROW_LANE, ROW_V, ROW_OBS = 0, 10, 22, 27


def build_qp(asm, x0, U, obs_sv, n_obs):
    """QP in the absolute variable x = U+ :  min 1/2 x'Px + q'x,  l <= A x <= u."""
    B = x0.shape[0]
    Jr, r, X, dX = asm["Jr"], asm["r"], asm["X"], asm["dX"]
    JW = Jr * (2.0 * W15)[None, :, None]
    Pm = np.einsum("bki,bkj->bij", JW, Jr) + np.eye(NV)[None]
    g = np.einsum("bki,bk->bi", JW, r) + U
    q = g - np.einsum("bij,bj->bi", Pm, U)
    A = np.zeros((B, M, NV))
    l = np.full((B, M), -INF)
    u = np.full((B, M), INF)
    A[:, 0:10] = np.eye(NV)
    l[:, 0:10] = np.tile(P.U_MIN, 5)
    u[:, 0:10] = np.tile(P.U_MAX, 5)
    const_viol = np.zeros(B)          # worst violation among rows U cannot influence
    row = ROW_LANE
    for j in range(1, 6):
        for al in ALPHAS:
            a = dX[:, j, 1] + al * dX[:, j, 2]
            val = X[:, j, 1] + al * X[:, j, 2]
            if j == 1:
                const_viol = np.maximum(const_viol, np.abs(val) - SLD)
                continue
            c0 = val - np.einsum("bi,bi->b", a, U)
            A[:, row] = a
            l[:, row] = -SLD - c0
            u[:, row] = SLD - c0
            row += 1
    v0, s0 = x0[:, 4], x0[:, 0]
    for j in range(1, 6):
        A[:, ROW_V + j - 1] = SV[j - 1]
        l[:, ROW_V + j - 1] = -v0
    for k in range(2):
        on = n_obs > k
        for j in range(1, 6):
            sobs = obs_sv[:, k, 0] + obs_sv[:, k, 1] * (j * H)
            base = sobs - s0 - j * H * v0
            r1 = ROW_OBS + 10 * k + 2 * (j - 1)
            if j == 1:
                const_viol = np.maximum(const_viol, np.where(on, P.OBS_SAFETY_DIST - base, 0.0))
            else:
                A[:, r1] = SS[j - 1]
                u[:, r1] = np.where(on, base - P.OBS_SAFETY_DIST, INF)
            A[:, r1 + 1] = SS[j - 1] + P.MAX_TIME_2_OBS * SV[j - 1]
            u[:, r1 + 1] = np.where(on, base - P.MAX_TIME_2_OBS * v0, INF)
    return dict(P=Pm, q=q, g=g, A=A, l=l, u=u, const_viol=const_viol)


# ----------------------------------------------------------------------------------------------
def admm(qp, x, z, y, rho, sigma=1e-6, alpha=1.6, iters=50, row_scale=None):
    """OSQP iteration with per-row step sizes rho (B,M).  State (x,z,y) in/out."""
    Pm, q, A, l, u = qp["P"], qp["q"], qp["A"], qp["l"], qp["u"]
    K = Pm + sigma * np.eye(NV)[None] + np.einsum("bmi,bm,bmj->bij", A, rho, A)
    Kinv = np.linalg.inv(K)
    for _ in range(iters):
        rhs = sigma * x - q + np.einsum("bmi,bm->bi", A, rho * z - y)
        xt = np.einsum("bij,bj->bi", Kinv, rhs)
        zt = np.einsum("bmi,bi->bm", A, xt)
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        znew = np.clip(zr + y / rho, l, u)
        y = y + rho * (zr - znew)
        z = znew
    return x, z, y


def residuals(qp, x, z, y):
    Pm, q, A = qp["P"], qp["q"], qp["A"]
    Ax = np.einsum("bmi,bi->bm", A, x)
    rp = np.max(np.abs(Ax - z), axis=1)
    rd = np.max(np.abs(np.einsum("bij,bj->bi", Pm, x) + q + np.einsum("bmi,bm->bi", A, y)), axis=1)
    return rp, rd


def polish(qp, x, y, act_tol=0.0, delta=1e-9, refine=3):
    """OSQP-style polish: guess active rows from the sign of y, solve the equality-constrained
    KKT system through the Schur complement of P with regularisation delta and iterative
    refinement.  Returns x_pol, y_pol (full length M), ok flag (primal feasible & dual signs)."""
    Pm, q, A, l, u = qp["P"], qp["q"], qp["A"], qp["l"], qp["u"]
    B = x.shape[0]
    xo = x.copy()
    yo = np.zeros_like(y)
    ok = np.zeros(B, dtype=bool)
    for b in range(B):
        lo = np.where(y[b] < -act_tol)[0]
        up = np.where(y[b] > act_tol)[0]
        idx = np.concatenate([lo, up])
        Aa = A[b, idx]
        ba = np.concatenate([l[b, lo], u[b, up]])
        na = len(idx)
        if na == 0:
            xs = np.linalg.solve(Pm[b], -q[b])
            ya = np.zeros(0)
        else:
            Kk = np.block([[Pm[b] + delta * np.eye(NV), Aa.T], [Aa, -delta * np.eye(na)]])
            Kt = np.block([[Pm[b], Aa.T], [Aa, np.zeros((na, na))]])
            rhs = np.concatenate([-q[b], ba])
            sol = np.linalg.solve(Kk, rhs)
            for _ in range(refine):
                sol = sol + np.linalg.solve(Kk, rhs - Kt @ sol)
            xs, ya = sol[:NV], sol[NV:]
        Ax = A[b] @ xs
        feas = np.all(Ax >= l[b] - 1e-9) and np.all(Ax <= u[b] + 1e-9)
        sign = np.all(ya[: len(lo)] <= 1e-9) and np.all(ya[len(lo):] >= -1e-9)
        xo[b] = xs
        yo[b, idx] = ya
        ok[b] = feas and sign
    return xo, yo, ok


# ----------------------------------------------------------------------------------------------
def row_rho(qp, rho0):
    """Per-row step size: rho0 / ||a_i||^2 (equivalent to normalising every row to unit norm)."""
    nrm2 = np.sum(qp["A"] ** 2, axis=2)
    return rho0 / np.maximum(nrm2, 1e-12)


def sqp_solve(tab, x0, obs_sv, n_obs, rounds=6, iters=60, rho0=1.0, alpha=1.6, sigma=1e-6,
              step_tol=1e-7, verbose=False, use_polish=True, trace=None):
    B = x0.shape[0]
    U = warm_start(tab, x0, obs_sv, n_obs)
    x = U.copy()
    z = None
    y = np.zeros((B, M))
    done = np.zeros(B, dtype=bool)
    nround = np.zeros(B, dtype=int)
    okflag = np.zeros(B, dtype=bool)
    for r in range(rounds):
        asm = assemble(tab, x0, U)
        qp = build_qp(asm, x0, U, obs_sv, n_obs)
        rho = row_rho(qp, rho0)
        if z is None:
            z = np.clip(np.einsum("bmi,bi->bm", qp["A"], x), qp["l"], qp["u"])
        x, z, y = admm(qp, x, z, y, rho, sigma=sigma, alpha=alpha, iters=iters)
        if use_polish:
            xp, yp, ok = polish(qp, x, y)
            xn = np.where(ok[:, None], xp, x)
        else:
            ok = np.zeros(B, dtype=bool)
            xn = x
        step = np.max(np.abs(xn - U), axis=1)
        upd = ~done
        U = np.where(upd[:, None], xn, U)
        okflag = np.where(upd, ok, okflag)
        nround += upd
        done |= (step < step_tol) & ok
        if trace is not None:
            trace.append(U.copy())
        if verbose:
            rp, rd = residuals(qp, x, z, y)
            print(f"round {r}: step max {step.max():.2e} med {np.median(step):.2e} ok {ok.mean():.3f} "
                  f"done {done.mean():.3f} rp {np.median(rp):.1e}/{rp.max():.1e} rd {np.median(rd):.1e}/{rd.max():.1e}")
    asm = assemble(tab, x0, U)
    return dict(U=U, cost=asm["cost"], X=asm["X"], rounds=nround, ok=okflag, done=done)


# ----------------------------------------------------------------------------------------------
def slsqp_exact(tab, x0, obs_sv, n_obs, U_start=None, ftol=1e-14, maxiter=300):
    """Development oracle: reference formulation, SciPy SLSQP with EXACT gradients/Jacobians from
    ``assemble`` (lower noise floor than the finite-difference oracle).  One problem."""
    from scipy.optimize import minimize
    x0b = np.asarray(x0, dtype=np.float64)[None]
    obs = [(float(obs_sv[k, 0]), float(obs_sv[k, 1])) for k in range(int(n_obs))]

    def f(U):
        a = assemble(tab, x0b, U[None])
        g = np.einsum("k,ki->i", 2.0 * W15 * a["r"][0], a["Jr"][0]) + U
        return float(a["cost"][0]), g

    def c(U):
        return P.constraint_values(tab, U, x0b[0], obs)

    def cj(U):
        a = assemble(tab, x0b, U[None])
        X, dX = a["X"][0], a["dX"][0]
        rows = []
        for j in range(1, 6):
            for al in ALPHAS:
                gr = dX[j, 1] + al * dX[j, 2]
                rows.append(-gr)
                rows.append(gr)
            for (so, vo) in obs:
                if X[j, 4] * P.MAX_TIME_2_OBS > P.OBS_SAFETY_DIST:
                    rows.append(-dX[j, 0] - P.MAX_TIME_2_OBS * dX[j, 4])
                else:
                    rows.append(-dX[j, 0])
            rows.append(dX[j, 4])
        return np.array(rows)

    U0 = warm_start(tab, x0b, obs_sv[None], np.array([n_obs]))[0] if U_start is None else U_start
    sol = minimize(f, U0, jac=True, method="SLSQP", bounds=P.bounds_list(),
                   constraints={"type": "ineq", "fun": c, "jac": cj},
                   options={"ftol": ftol, "maxiter": maxiter, "disp": False})
    return sol


# ----------------------------------------------------------------------------------------------
# Algorithm v2 (the one the CUDA kernel implements): sigma = 0, single-vector ADMM state
# v = z_relaxed + y/rho, per-row rho re-selected from the detected active set every segment.
NRM2_FLOOR = 1e-2


def seg_rho(qp, act, rho_lo, rho_hi):
    nrm2 = np.maximum(np.sum(qp["A"] ** 2, axis=2), NRM2_FLOOR)
    return np.where(act, rho_hi, rho_lo) / nrm2


def admm_v(qp, z, y, rho, alpha, iters):
    """sigma=0 OSQP iteration.  Returns x (last x~), z, y, rp, rd (inf-norm residuals of the last
    iterate: rp = |A x - z|, rd = |P x + q + A'y| evaluated without P), dy (last y increment)."""
    Pm, q, A, l, u = qp["P"], qp["q"], qp["A"], qp["l"], qp["u"]
    K = Pm + np.einsum("bmi,bm,bmj->bij", A, rho, A)
    Kinv = np.linalg.inv(K)
    for _ in range(iters):
        w = rho * z - y
        x = np.einsum("bij,bj->bi", Kinv, np.einsum("bmi,bm->bi", A, w) - q)
        zt = np.einsum("bmi,bi->bm", A, x)
        zr = alpha * zt + (1 - alpha) * z
        zn = np.clip(zr + y / rho, l, u)
        dy = rho * (zr - zn)
        # residuals of (x, zn, y+dy):  P x + q = A'(rho z - y) - A' rho zt
        rd = np.max(np.abs(np.einsum("bmi,bm->bi", A, rho * (z - zt) + dy)), axis=1)
        rp = np.max(np.abs(zt - zn), axis=1)
        y = y + dy
        z = zn
    return x, z, y, rp, rd, dy


def sqp_solve_v2(tab, x0, obs_sv, n_obs, max_rounds=10, seg_iters=10, max_segs=8, rho_mid=1.0,
                 rho_lo=0.1, rho_hi=1e4, alpha=1.6, eps_p=1e-9, eps_d=1e-8, step_tol=1e-7,
                 verbose=False):
    B = x0.shape[0]
    U = warm_start(tab, x0, obs_sv, n_obs)
    z = None
    y = np.zeros((B, M))
    done = np.zeros(B, dtype=bool)
    rounds = np.zeros(B, dtype=int)
    iters_tot = np.zeros(B, dtype=int)
    qp_ok = np.zeros(B, dtype=bool)
    for r in range(max_rounds):
        asm = assemble(tab, x0, U)
        qp = build_qp(asm, x0, U, obs_sv, n_obs)
        if z is None:
            z = np.clip(np.einsum("bmi,bi->bm", qp["A"], U), qp["l"], qp["u"])
        conv = done.copy()
        x = U.copy()
        for sgm in range(max_segs):
            if r == 0 and sgm == 0:
                rho = seg_rho(qp, np.zeros((B, M), bool), rho_mid, rho_mid)
            else:
                rho = seg_rho(qp, np.abs(y) > 1e-10, rho_lo, rho_hi)
            xn, zn, yn, rp, rd, dy = admm_v(qp, z, y, rho, alpha, seg_iters)
            run = ~conv
            x = np.where(run[:, None], xn, x)
            z = np.where(run[:, None], zn, z)
            y = np.where(run[:, None], yn, y)
            iters_tot += run * seg_iters
            conv |= (rp <= eps_p) & (rd <= eps_d)
            if conv.all():
                break
        step = np.max(np.abs(x - U), axis=1)
        upd = ~done
        U = np.where(upd[:, None], x, U)
        qp_ok = np.where(upd, conv, qp_ok)
        rounds += upd
        done |= (step < step_tol) & conv
        if verbose:
            print(f"round {r}: step max {step.max():.2e} med {np.median(step):.2e} qp_conv {conv.mean():.3f} "
                  f"done {done.mean():.3f} iters mean {iters_tot.mean():.1f}")
        if done.all():
            break
    asm = assemble(tab, x0, U)
    return dict(U=U, cost=asm["cost"], X=asm["X"], rounds=rounds, iters=iters_tot, done=done, qp_ok=qp_ok)
