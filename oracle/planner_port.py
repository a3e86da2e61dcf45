"""CPU oracle for the planner function evaluation (north_star item (c)).  TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's offline Hermite-Simpson NLP functions (file:line relative to the upstream
repo root):
    dynamics                      trajectory_planning.py:49-89
    unpack / pack                 trajectory_planning.py:91-126
    cost                          trajectory_planning.py:128-170
    dynamics_constraints closure  trajectory_planning.py:183-208
    inequality closures           trajectory_planning.py:249-347
It is the checker for ``mpcb_hs_eval`` / ``mpcb_hs_nodes``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs may import it; the product package never does.

Parity status: the reference ships no tests for this path ("parity unpinned" by the reference itself).  The port is
pinned against the unmodified reference imported in the build container through a stub ``path_planning`` module
(the real one needs pymap3d, an API key and network): ``tools/make_golden_planner.py`` writes
``tests/golden/planner_traj{1,2,3}.npz``; ``tests/test_planner_oracle.py`` checks this port against them.

Third-party arithmetic: ``k_ref_fun`` / ``v_max_fun`` of the reference come from GraphHopper data that is not in the
repository.  The synthetic substitute defined in SURVEY.md 8(d) is used everywhere:
``k_ref_fun = TrajectoryLoader.interp_k`` (scipy interp1d linear with extrapolation, restated in ``RefTable._lin``),
``v_max_fun = const``, ``v_min_fun = 0``.

The functions are written so that they also accept complex arguments: the Jacobian oracle is a complex-step
derivative of these very functions (exact to rounding, unlike finite differences).
"""
from __future__ import annotations

import numpy as np

DT = 0.3                                   # trajectory_planning.py:514
W_Y, W_S, W_U, W_SLACK = 10.0, 10.0, 0.1, 100.0   # :14
U_MIN = np.array([-0.6, -5.0])             # :36
U_MAX = np.array([0.6, 4.0])               # :37
K_MIN, K_MAX = -0.8, 0.8                   # :44-45
A_MAX = 6.0                                # :48


def kref_linear(tab, s):
    """TrajectoryLoader.interp_k(s): linear interpolation / extrapolation of the curvature column
    (trajectory_loader.py:69 + scipy _call_linear).  Accepts complex s (segment chosen by the real part)."""
    sr = np.real(s)
    S = tab.s
    i = int(np.clip(np.searchsorted(S, sr, side="left"), 1, tab.K - 1))
    lo = i - 1
    x_lo, x_hi = S[lo], S[i]
    y_lo, y_hi = tab.X[lo, 3], tab.X[i, 3]
    return ((s - x_lo) / (x_hi - x_lo)) * y_hi + ((x_hi - s) / (x_hi - x_lo)) * y_lo


def dynamics(x, u, k_ref):
    """trajectory_planning.py:49-89"""
    s, d, o, k, v = x
    u1, u2 = u
    denom = 1 - d * k_ref
    if abs(denom) < 1e-4:                                   # :74-77
        denom = 1e-4 * np.sign(np.real(denom)) if denom != 0 else 1e-4
    s_dot = (v * np.cos(o)) / denom
    d_dot = v * np.sin(o)
    o_dot = v * k - s_dot * k_ref
    return np.array([s_dot, d_dot, o_dot, u1, u2])


def unpack(z, N):
    """:91-115"""
    X = z[: 5 * (N + 1)].reshape(N + 1, 5)
    U = z[5 * (N + 1): 5 * (N + 1) + 2 * N].reshape(N, 2)
    S = z[5 * (N + 1) + 2 * N:]
    return X, U, S


def pack(X, U, S):
    """:117-126"""
    return np.concatenate([np.ravel(X), np.ravel(U), np.ravel(S)])


def hs_defect(tab, x_k, x_next, u_k, dt=DT, simpson_sign=-1):
    """Closure dynamics_constraints (:183-208) for one interval.  simpson_sign=-1 is the committed code (:198)."""
    f_k = dynamics(x_k, u_k, kref_linear(tab, x_k[0]))
    f_next = dynamics(x_next, u_k, kref_linear(tab, x_next[0]))
    x_mid = 0.5 * (x_k + x_next) + (dt / 8.0) * (f_k - f_next)
    f_mid = dynamics(x_mid, u_k, kref_linear(tab, x_mid[0]))
    simpson = (dt / 6.0) * (f_k + 4 * f_mid + f_next)
    x_pred = x_k - simpson if simpson_sign < 0 else x_k + simpson
    return x_next - x_pred


def defects(tab, z, N, dt=DT, simpson_sign=-1):
    X, U, _ = unpack(z, N)
    return np.array([hs_defect(tab, X[k], X[k + 1], U[k], dt, simpson_sign) for k in range(N)])


def hs_defect_jac(tab, x_k, x_next, u_k, dt=DT, simpson_sign=-1, h=1e-30):
    """d defect / d (x_k, x_next, u_k) (5 x 12) by complex-step differentiation of hs_defect."""
    y0 = np.concatenate([x_k, x_next, u_k]).astype(np.complex128)
    J = np.zeros((5, 12))
    for j in range(12):
        y = y0.copy()
        y[j] += 1j * h
        J[:, j] = np.imag(hs_defect(tab, y[0:5], y[5:10], y[10:12], dt, simpson_sign)) / h
    return J


def hs_lagrangian_hess(tab, x_k, x_next, u_k, lam, dt=DT, simpson_sign=-1, eps=1e-6):
    """d^2 (lam . defect) / d y^2 (12 x 12) by central differences of the complex-step gradient."""
    y0 = np.concatenate([x_k, x_next, u_k])

    def grad(y):
        return hs_defect_jac(tab, y[0:5], y[5:10], y[10:12], dt, simpson_sign).T @ lam

    H = np.zeros((12, 12))
    for j in range(12):
        e = np.zeros(12)
        e[j] = eps
        H[:, j] = (grad(y0 + e) - grad(y0 - e)) / (2 * eps)
    return 0.5 * (H + H.T)


def cost(z, N, x0, s_total):
    """:128-170"""
    X, U, S = unpack(z, N)
    c = 0.0
    for k in range(N):
        s_k, d_k, o_k, k_k, v_k = X[k]
        y_k = np.array([d_k, o_k])
        term1 = W_Y * (y_k @ y_k)
        denom = max(1, s_total - x0[0])
        term2 = W_S * ((s_total - s_k) / denom) ** 2
        term3 = W_U * (U[k] @ U[k])
        term4 = W_SLACK * (S[k] ** 2)
        c += term1 + term2 + term3 + term4
    return c


def stage_costs(z, N, x0, s_total):
    X, U, S = unpack(z, N)
    denom = max(1, s_total - x0[0])
    out = np.zeros(N)
    for k in range(N):
        out[k] = (W_Y * (X[k, 1] * X[k, 1] + X[k, 2] * X[k, 2]) + W_S * ((s_total - X[k, 0]) / denom) ** 2
                  + W_U * (U[k] @ U[k]) + W_SLACK * (S[k] ** 2))
    return out


def node_rows(z, N, v_min, v_max):
    """Inequality closures per node, :249-307, in the order speed_min, speed_max, lateral_max, lateral_min,
    k_min, k_max.  v_min / v_max: scalars or arrays over nodes (the reference calls v_*_fun(s_k))."""
    X, _, S = unpack(z, N)
    vmin = np.broadcast_to(v_min, (N + 1,))
    vmax = np.broadcast_to(v_max, (N + 1,))
    out = np.zeros((N + 1, 6))
    for k in range(N + 1):
        slack = S[k] if k < N else 0
        k_k, v_k = X[k, 3], X[k, 4]
        out[k] = [(v_k + slack) - vmin[k], vmax[k] - (v_k + slack), A_MAX - (k_k * v_k ** 2), A_MAX + (k_k * v_k ** 2),
                  k_k - K_MIN, K_MAX - k_k]
    return out


def ctrl_rows(z, N):
    """:310-347: u1_min, u1_max, u2_min, u2_max, non_negative_slack per interval."""
    _, U, S = unpack(z, N)
    out = np.zeros((N, 5))
    for k in range(N):
        out[k] = [U[k, 0] - U_MIN[0], U_MAX[0] - U[k, 0], U[k, 1] - U_MIN[1], U_MAX[1] - U[k, 1], S[k]]
    return out


class OracleEvaluator:
    """CPU stand-in for PlannerEvaluator.evaluate_host built from the functions above (values) and complex-step
    derivatives.  TEST INFRASTRUCTURE: lets tests drive ``planner_driver`` without a GPU and check the GPU evaluator
    against it."""

    def __init__(self, tab, N, dt=DT, simpson_sign=-1, s_total=None, v_min=0.0, v_max=None):
        self.tab, self.N, self.dt, self.sign = tab, int(N), dt, simpson_sign
        self.s_total = float(tab.s_max if s_total is None else s_total)
        self.v_min = v_min
        self.v_max = float(tab.X[:, 4].max() if v_max is None else v_max)

    def evaluate_host(self, z, lam=None, s0=None, want_jac=True, want_hess=False):
        z = np.atleast_2d(np.asarray(z, dtype=np.float64))
        C, N = z.shape[0], self.N
        s0 = z[:, 0] if s0 is None else np.asarray(s0, dtype=np.float64).reshape(C)
        out = dict(defect=np.zeros((C, N, 5)), node_rows=np.zeros((C, N + 1, 6)), ctrl_rows=np.zeros((C, N, 5)),
                   cost_terms=np.zeros((C, N)), cost=np.zeros(C), cost_grad=np.zeros((C, 8 * N + 5)))
        if want_jac:
            out["jac"] = np.zeros((C, N, 5, 12))
        for c in range(C):
            X, U, S = unpack(z[c], N)
            out["defect"][c] = defects(self.tab, z[c], N, self.dt, self.sign)
            out["node_rows"][c] = node_rows(z[c], N, self.v_min, self.v_max)
            out["ctrl_rows"][c] = ctrl_rows(z[c], N)
            x0 = np.array([s0[c], 0, 0, 0, 0.0])
            out["cost_terms"][c] = stage_costs(z[c], N, x0, self.s_total)
            out["cost"][c] = cost(z[c], N, x0, self.s_total)
            denom = max(1, self.s_total - s0[c])
            g = out["cost_grad"][c]
            for k in range(N):
                g[5 * k] = -2.0 * W_S * ((self.s_total - X[k, 0]) / denom) / denom
                g[5 * k + 1] = 2.0 * W_Y * X[k, 1]
                g[5 * k + 2] = 2.0 * W_Y * X[k, 2]
                g[5 * (N + 1) + 2 * k] = 2.0 * W_U * U[k, 0]
                g[5 * (N + 1) + 2 * k + 1] = 2.0 * W_U * U[k, 1]
                g[7 * N + 5 + k] = 2.0 * W_SLACK * S[k]
            if want_jac:
                for k in range(N):
                    out["jac"][c, k] = hs_defect_jac(self.tab, X[k], X[k + 1], U[k], self.dt, self.sign)
        return out
