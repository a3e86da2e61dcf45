"""Host driver of the offline planner around the GPU function evaluator.

Mirrors ``TrajectoryOptimizer.optimize`` (trajectory_planning.py:351-390: initial guess, SLSQP, maxiter 500, ftol 1e-4)
and the receding-horizon loop of ``optimize_full_trajectory`` (:491-554: 20 m chunks, horizon from the mean speed limit,
dt = 0.3, commit the first N/2 steps).  The outer optimisation stays on the host by design (north_star item (c)); what
changes is that every function value and every derivative SLSQP asks for comes from ONE batched GPU evaluation of all
collocation intervals (``PlannerEvaluator.evaluate_host``) instead of ~15,000 Python closure calls per finite-difference
Jacobian.

``evaluator`` is anything with ``N`` and ``evaluate_host(z, s0=..., want_jac=True) -> dict(defect, jac, node_rows,
ctrl_rows, cost, cost_grad)``; the tests plug a CPU stand-in in here to check the driver logic without a GPU.
"""
import numpy as np
from scipy.optimize import minimize


class ChunkProblem:
    """Constraint stack of one chunk in the reference's order (trajectory_planning.py:172-349), with analytic Jacobians
    assembled from the evaluator's per-interval / per-node blocks."""

    def __init__(self, evaluator, x0, s_target, is_final_chunk):
        self.ev = evaluator
        self.N = N = evaluator.N
        self.nz = 8 * N + 5
        self.x0 = np.asarray(x0, dtype=np.float64)
        self.s_target = float(s_target)
        self.final = bool(is_final_chunk)
        self._key = None
        self._val = None
        self.n_eval = 0
        # constant Jacobian pieces
        self._J_init = np.zeros((5, self.nz))
        self._J_init[np.arange(5), np.arange(5)] = 1.0
        self._J_term_s = np.zeros(self.nz)
        self._J_term_s[5 * N] = 1.0
        self._J_term_v = np.zeros(self.nz)
        self._J_term_v[5 * N + 4] = 1.0

    def _eval(self, z):
        key = z.tobytes()
        if key != self._key:
            self._val = self.ev.evaluate_host(z, s0=np.array([self.x0[0]]), want_jac=True)
            self._key = key
            self.n_eval += 1
        return self._val

    # ---- objective (:128-170) ---------------------------------------------------------------------------
    def cost(self, z):
        return float(self._eval(z)["cost"][0])

    def cost_grad(self, z):
        return np.array(self._eval(z)["cost_grad"][0])

    # ---- equalities: N defect closures, initial_x0, terminal (final chunk) --------------------------------
    def eq(self, z):
        r = self._eval(z)
        X0 = z[:5] - self.x0
        parts = [r["defect"][0].ravel(), X0]
        if self.final:
            parts.append(np.array([z[5 * self.N] - self.s_target, z[5 * self.N + 4]]))
        return np.concatenate(parts)

    def eq_jac(self, z):
        N = self.N
        J = self._eval(z)["jac"][0]                     # [N][5][12] w.r.t. (x_k, x_{k+1}, u_k)
        rows = np.zeros((5 * N, self.nz))
        for k in range(N):
            rows[5 * k:5 * k + 5, 5 * k:5 * k + 10] = J[k][:, :10]
            rows[5 * k:5 * k + 5, 5 * (N + 1) + 2 * k:5 * (N + 1) + 2 * k + 2] = J[k][:, 10:12]
        parts = [rows, self._J_init]
        if self.final:
            parts.append(np.vstack([self._J_term_s, self._J_term_v]))
        return np.vstack(parts)

    # ---- inequalities: terminal_s (intermediate chunks), node rows, control rows ---------------------------
    def ineq(self, z):
        r = self._eval(z)
        parts = []
        if not self.final:
            parts.append(np.array([z[5 * self.N] - self.s_target / 2]))
        nr = r["node_rows"][0]
        parts += [nr[:, 0:4].ravel(), nr[:, 4:6].ravel(), r["ctrl_rows"][0].ravel()]
        return np.concatenate(parts)

    def ineq_jac(self, z):
        N, nz = self.N, self.nz
        X = z[:5 * (N + 1)].reshape(N + 1, 5)
        rows = []
        if not self.final:
            rows.append(self._J_term_s[None])
        A = np.zeros((4 * (N + 1), nz))
        Bm = np.zeros((2 * (N + 1), nz))
        for k in range(N + 1):
            kk, vv = X[k, 3], X[k, 4]
            sl = 7 * N + 5 + k if k < N else None
            A[4 * k, 5 * k + 4] = 1.0                   # (v + slack) - v_min   (constant limits: d/ds = 0)
            A[4 * k + 1, 5 * k + 4] = -1.0              # v_max - (v + slack)
            if sl is not None:
                A[4 * k, sl] = 1.0
                A[4 * k + 1, sl] = -1.0
            A[4 * k + 2, 5 * k + 3] = -vv * vv          # a_max - k v^2
            A[4 * k + 2, 5 * k + 4] = -2.0 * kk * vv
            A[4 * k + 3, 5 * k + 3] = vv * vv           # a_max + k v^2
            A[4 * k + 3, 5 * k + 4] = 2.0 * kk * vv
            Bm[2 * k, 5 * k + 3] = 1.0                  # k - k_min
            Bm[2 * k + 1, 5 * k + 3] = -1.0             # k_max - k
        Cm = np.zeros((5 * N, nz))
        for k in range(N):
            u = 5 * (N + 1) + 2 * k
            Cm[5 * k, u] = 1.0
            Cm[5 * k + 1, u] = -1.0
            Cm[5 * k + 2, u + 1] = 1.0
            Cm[5 * k + 3, u + 1] = -1.0
            Cm[5 * k + 4, 7 * N + 5 + k] = 1.0
        rows += [A, Bm, Cm]
        return np.vstack(rows)


def initial_guess(N, x0, s_target, is_final_chunk):
    """trajectory_planning.py:358-376"""
    X = np.zeros((N + 1, 5))
    X[:, 0] = np.linspace(x0[0], s_target, N + 1)
    X[:, 4] = np.linspace(x0[4], 0.0, N + 1) if is_final_chunk else x0[4]
    return np.concatenate([X.ravel(), np.zeros(2 * N), np.zeros(N)])


def optimize_chunk(evaluator, x0, s_target, is_final_chunk, maxiter=500, ftol=1e-4, z0=None):
    """``TrajectoryOptimizer.optimize`` (:351-390) with GPU-evaluated functions and analytic derivatives.
    Returns (X, U, S, scipy result, number of batched evaluations)."""
    N = evaluator.N
    prob = ChunkProblem(evaluator, x0, s_target, is_final_chunk)
    if z0 is None:
        z0 = initial_guess(N, np.asarray(x0, dtype=np.float64), s_target, is_final_chunk)
    cons = [{"type": "eq", "fun": prob.eq, "jac": prob.eq_jac}, {"type": "ineq", "fun": prob.ineq, "jac": prob.ineq_jac}]
    sol = minimize(prob.cost, z0, jac=prob.cost_grad, method="SLSQP", constraints=cons,
                   options={"maxiter": maxiter, "ftol": ftol, "disp": False})
    z = sol.x
    X = z[:5 * (N + 1)].reshape(N + 1, 5)
    U = z[5 * (N + 1):5 * (N + 1) + 2 * N].reshape(N, 2)
    S = z[5 * (N + 1) + 2 * N:]
    return X, U, S, sol, prob.n_eval


def optimize_full_trajectory(make_evaluator, s_total, v_max, max_chunk_size=20.0, dt=0.3, max_chunks=100000):
    """Receding-horizon loop of trajectory_planning.py:479-554 for a route described by the evaluator's own k_ref table
    and a constant speed limit ``v_max`` (the GraphHopper speed-limit arrays are not in the repository).
    ``make_evaluator(N, s_total)`` returns an evaluator for chunks of N intervals.  Returns X, U, S, per-chunk log."""
    x0 = np.array([0.0, 0.0, 0.0, 0.0, 0.0])                       # :486
    Xf, Uf, Sf, log = [], [], [], []
    remaining = s_total
    while remaining > 0.1 and len(log) < max_chunks:                # :491
        if remaining < max_chunk_size * 2:                          # :494-499
            chunk, final = remaining, True
        else:
            chunk, final = max_chunk_size, False
        s_target = x0[0] + chunk
        horizon = (chunk / v_max) * 2.0                             # :507-512 (avg speed limit, safety buffer 2)
        N = int(np.ceil(horizon / dt))
        ev = make_evaluator(N, s_total)
        X, U, S, sol, n_eval = optimize_chunk(ev, x0, s_target, final)
        log.append(dict(N=N, final=final, status=int(sol.status), nit=int(sol.nit), n_eval=n_eval, cost=float(sol.fun)))
        if not final:                                               # :523-541
            c = int(N / 2)
            Xs, Us, Ss = X[:c + 1], U[:c], S[:c]
            Xf.append(Xs if not Xf else Xs[1:])
            Uf.append(Us)
            Sf.append(Ss)
        else:                                                       # :542-545
            Xf.append(X[1:])
            Uf.append(U)
            Sf.append(S)
        x0 = Xf[-1][-1].copy()                                      # :548
        remaining = s_total - x0[0]
    return np.concatenate(Xf), np.concatenate(Uf), np.concatenate(Sf), log


def reference_trajectory_verdicts(X, U, S, s_total, u_min=(-0.6, -5.0), u_max=(0.6, 4.0)):
    """The items of ``reference_trajectory_check`` (sanity_checks.py:3-75) as a dict of booleans (the reference prints
    them and returns None; its control-limit test uses ``and`` where ``or`` is meant, :48/:54 -- reproduced as is)."""
    return dict(
        destination=abs(X[-1, 0] - s_total) <= 0.5,
        full_stop=abs(X[-1, 4]) <= 0.1,
        non_negative_velocity=not (np.min(X[:, 4]) < -0.1),
        u1_limits=not (np.min(U[:, 0]) < u_min[0] - 0.1 and np.max(U[:, 0]) > u_max[0] + 0.1),
        u2_limits=not (np.min(U[:, 1]) < u_min[1] - 0.1 and np.max(U[:, 1]) > u_max[1] + 0.1),
        lateral=not (np.max(np.abs(X[:, 1])) > 1.5),
        slack=not (np.max(np.abs(S)) > 0.1),
    )
