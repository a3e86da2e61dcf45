"""Test helper: optimality and feasibility of returned controls judged with the ORACLE's functions only.

``reference_rows`` evaluates the reference's constraints_wrapper (trajectory_tracking.py:164-209) for a whole batch from
the oracle model's rollout (oracle/sqp_admm_model.assemble, itself pinned to the reference's predict/cost by
tests/test_algorithm_model.py) together with the exact gradient of every row; ``kkt_residual`` then asks, per problem,
whether non-negative multipliers on the active rows and bounds reproduce the cost gradient (a non-negative least-squares
fit) -- first-order optimality of the NONLINEAR problem, including the problems that sit on constraints.
Nothing here reads a value the device computed except the controls under test."""
import numpy as np
from scipy.optimize import nnls

from oracle import sqp_admm_model as A
from oracle import tracker_port as P


def reference_rows(tab, x0, U, obs_sv, n_obs, with_kinks=False):
    """c [B,45] (NaN beyond 5*(7+n_obs)) in the reference's row order, G [B,45,10] = dc/dU, asm (the rollout).
    The obstacle row gap - max(5, 1.5 v) (trajectory_tracking.py:201) has a kink at 1.5 v = 5: G holds the gradient of
    the branch that is larger; with_kinks=True additionally returns G2 [B,45,10], the gradient of the OTHER branch, and
    kink [B,45] = how far the two branches are apart (NaN for rows without a kink) -- at a kink both are subgradients."""
    B = len(x0)
    asm = A.assemble(tab, x0, U)
    X, dX = asm["X"], asm["dX"]
    c = np.full((B, 45), np.nan)
    G = np.zeros((B, 45, 10))
    G2 = np.zeros((B, 45, 10))
    kink = np.full((B, 45), np.nan)
    row = np.zeros(B, dtype=np.int64)
    ar = np.arange(B)
    for j in range(1, 6):
        s, d, o, v = X[:, j, 0], X[:, j, 1], X[:, j, 2], X[:, j, 4]
        ds, dd, do, dv = dX[:, j, 0], dX[:, j, 1], dX[:, j, 2], dX[:, j, 4]
        for al in A.ALPHAS:
            val, g = d + al * o, dd + al * do
            c[ar, row] = A.SLD - val
            G[ar, row] = -g
            c[ar, row + 1] = val + A.SLD
            G[ar, row + 1] = g
            row = row + 2
        for k in range(2):
            on = n_obs > k
            s_obs = obs_sv[:, k, 0] + obs_sv[:, k, 1] * (j * A.H)
            tg = P.MAX_TIME_2_OBS * v
            safe = np.maximum(P.OBS_SAFETY_DIST, tg)
            val = (s_obs - s) - safe
            g = -ds - np.where((tg > P.OBS_SAFETY_DIST)[:, None], P.MAX_TIME_2_OBS * dv, 0.0)
            idx = ar[on]
            c[idx, row[on]] = val[on]
            G[idx, row[on]] = g[on]
            g2 = -ds - np.where((tg > P.OBS_SAFETY_DIST)[:, None], 0.0, P.MAX_TIME_2_OBS * dv)
            G2[idx, row[on]] = g2[on]
            kink[idx, row[on]] = np.abs(tg - P.OBS_SAFETY_DIST)[on]
            row = row + on.astype(np.int64)
        c[ar, row] = v
        G[ar, row] = dv
        row = row + 1
    assert np.array_equal(row, 5 * (7 + n_obs))
    if with_kinks:
        return c, G, asm, G2, kink
    return c, G, asm


def cost_gradient(asm, U):
    return np.einsum("bk,bki->bi", 2.0 * A.W15 * asm["r"], asm["Jr"]) + U


def kkt_residual(c, G, g, U, act_tol=1e-4, G2=None, kink=None):
    """Per problem: min over lambda >= 0 of | grad J - sum_r lambda_r grad c_r - mu_lo + mu_hi |_inf over the rows and
    bounds within act_tol of active (a slightly generous candidate set only makes the fit easier, never wrong: any
    KKT point of the nonlinear problem has multipliers supported on its active set).  Returns res [B], n_active [B]."""
    B = len(U)
    lb, ub = np.tile(P.U_MIN, 5), np.tile(P.U_MAX, 5)
    res = np.zeros(B)
    nact = np.zeros(B, dtype=np.int64)
    for b in range(B):
        cols = []
        rows = np.where(c[b] <= act_tol)[0]          # NaN compares False
        for r in rows:
            cols.append(G[b, r])
            if kink is not None and kink[b, r] <= act_tol:     # on the kink of max(5, 1.5 v): both branches count
                cols.append(G2[b, r])
        for i in np.where(U[b] - lb <= act_tol)[0]:
            e = np.zeros(10); e[i] = 1.0
            cols.append(e)
        for i in np.where(ub - U[b] <= act_tol)[0]:
            e = np.zeros(10); e[i] = -1.0
            cols.append(e)
        nact[b] = len(cols)
        if not cols:
            res[b] = np.max(np.abs(g[b]))
            continue
        M = np.array(cols).T                          # 10 x n_active
        lam, _ = nnls(M, g[b])
        res[b] = np.max(np.abs(g[b] - M @ lam))
    return res, nact
