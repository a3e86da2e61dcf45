"""CPU oracle for the tracking-MPC hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy/scipy restatement of the reference's per-timestep tracking
MPC (reference files cited per function, relative to the upstream repo root).  It is
the *checker* for the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package never imports anything from ``oracle/`` and has no CPU fallback.

Parity status: the reference ships no tests and no golden vectors for this path
("parity unpinned" by the reference itself).  The port is therefore
pinned against outputs of the unmodified reference run in the build container:
``tools/make_golden.py`` imports the reference from ``/root/reference`` and writes
``tests/golden/*.npz``; ``tests/test_oracle_port.py`` checks this port against those
vectors and against SURVEY.md Appendix B's known-answer values.

Third-party arithmetic on the path (not under the reference tree):
  * scipy.interpolate.interp1d(kind='linear', fill_value='extrapolate') -- scipy is
    unpinned in the reference's requirements.txt:4; golden vectors were produced with
    scipy 1.18.1 / numpy 2.3.5.  Its published algorithm (``_call_linear``) is restated
    in ``RefTable._lin``.
  * scipy.optimize.minimize(method='SLSQP') -- called here exactly as the reference
    calls it (trajectory_tracking.py:254-256) for the as-shipped oracle, and at tight
    tolerance for the converged oracle.
"""
from __future__ import annotations

import json
import time

import numpy as np
from scipy.optimize import minimize

# --------------------------------------------------------------------------------------
# Parameter set -- trajectory_tracking.py:12-47
# --------------------------------------------------------------------------------------
DT = 0.2
N = 5
U_MIN = np.array([-0.6, -5.0])
U_MAX = np.array([0.6, 4.0])
VEHICLE_RADIUS = 1.0
W_D, W_O, W_V, W_U1, W_U2 = 10.0, 10.0, 5.0, 0.5, 0.5
OBS_SAFETY_DIST = 5.0
MAX_TIME_2_OBS = 1.5
WHEELBASE = 2.8
LANE_WIDTH = 3.0
SAFE_LANE_MARGIN = 0.1
BRAKE_LOOKAHEAD = 40.0   # trajectory_tracking.py:233
BRAKE_GUESS = -2.0       # trajectory_tracking.py:241


class RefTable:
    """Reference-signal table.  Follows trajectory_loader.py:13-30, :64-102.

    ``X`` is (K,5) rows [s,d,o,k,v]; ``U`` is (K-1,2) rows [u1,u2].
    """

    def __init__(self, X, U):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.U = np.ascontiguousarray(U, dtype=np.float64)
        s = self.X[:, 0].copy()
        # strict-monotone repair, trajectory_loader.py:27-30
        for i in range(1, len(s)):
            if s[i] <= s[i - 1]:
                s[i] = s[i - 1] + 1e-5
        self.s = s
        self.K = len(s)
        # control interpolators see only the first min(K, len(U)) knots, :73-77
        self.Ku = min(self.K, len(self.U))
        self.s_max = float(s[-1])   # :84

    @classmethod
    def from_json(cls, path):
        with open(path, "r") as f:
            data = json.load(f)
        return cls(np.array(data["X"]), np.array(data["U"]))

    @classmethod
    def from_npz(cls, path):
        z = np.load(path)
        return cls(z["X"], z["U"])

    @staticmethod
    def _lin(xs, ys, x):
        """scipy interp1d._call_linear (scipy/interpolate/_interpolate.py:491-517)."""
        i = int(np.searchsorted(xs, x, side="left"))
        i = min(max(i, 1), len(xs) - 1)
        lo = i - 1
        x_lo, x_hi = xs[lo], xs[i]
        return ((x - x_lo) / (x_hi - x_lo)) * ys[i] + ((x_hi - x) / (x_hi - x_lo)) * ys[lo]

    def get_state(self, s):
        """trajectory_loader.py:86-93."""
        if s >= self.s_max:
            return self.X[-1]
        return np.array([s,
                         self._lin(self.s, self.X[:, 1], s),
                         self._lin(self.s, self.X[:, 2], s),
                         self._lin(self.s, self.X[:, 3], s),
                         self._lin(self.s, self.X[:, 4], s)])

    def get_control(self, s):
        """trajectory_loader.py:95-102."""
        if s >= self.s_max:
            return np.array([0.0, 0.0])
        su = self.s[: self.Ku]
        return np.array([self._lin(su, self.U[: self.Ku, 0], s),
                         self._lin(su, self.U[: self.Ku, 1], s)])


# --------------------------------------------------------------------------------------
# Model functions
# --------------------------------------------------------------------------------------
def dynamics(x, u, k_ref):
    """trajectory_tracking.py:50-67."""
    return np.array([x[4], x[4] * x[2], x[4] * (x[3] - k_ref), u[0], u[1]])


def predict(tab: RefTable, x0, U_flat):
    """Explicit-Euler rollout, trajectory_tracking.py:87-114.  Returns (N+1,5)."""
    U = np.asarray(U_flat, dtype=np.float64).reshape(N, 2)
    X = np.zeros((N + 1, 5))
    X[0] = x0
    cur = np.array(x0, dtype=np.float64)
    for j in range(N):
        k_ref = tab.get_state(cur[0])[3]
        cur = cur + DT * dynamics(cur, U[j], k_ref)
        X[j + 1] = cur
    return X


def cost(tab: RefTable, U_flat, x0):
    """trajectory_tracking.py:116-152 (same accumulation order)."""
    U = np.asarray(U_flat, dtype=np.float64).reshape(N, 2)
    X = predict(tab, x0, U_flat)
    c = 0.0
    for j in range(1, N + 1):
        s, d, o, _k, v = X[j]
        r = tab.get_state(s)
        c += W_D * (d - r[1]) ** 2
        c += W_O * (o - r[2]) ** 2
        c += W_V * (v - r[4]) ** 2
    for j in range(N):
        c += W_U1 * U[j, 0] ** 2
        c += W_U2 * U[j, 1] ** 2
    return c


def constraint_values(tab: RefTable, U_flat, x0, obstacles):
    """constraints_wrapper, trajectory_tracking.py:164-209.  obstacles: list of (s, v)."""
    X = predict(tab, x0, U_flat)
    sld = LANE_WIDTH / 2.0 - VEHICLE_RADIUS - SAFE_LANE_MARGIN
    out = []
    for j in range(1, N + 1):
        s, d, o, v = X[j, 0], X[j, 1], X[j, 2], X[j, 4]
        out.append(sld - d)
        out.append(d + sld)
        vf = d + (WHEELBASE / 2.0) * o
        out.append(sld - vf)
        out.append(vf + sld)
        vl = d + WHEELBASE * o
        out.append(sld - vl)
        out.append(vl + sld)
        for (so, vo) in obstacles:
            s_obs = so + vo * (j * DT)
            gap = s_obs - s
            safe = max(OBS_SAFETY_DIST, v * MAX_TIME_2_OBS)
            out.append(gap - safe)
        out.append(v)
    return np.array(out)


def warm_start(tab: RefTable, x0, obstacles):
    """trajectory_tracking.py:223-246.  Returns flat (2N,) initial guess (unclipped)."""
    s_cur = x0[0]
    v_cur = x0[4]
    brake = False
    g = []
    for _ in range(N):
        for (so, _vo) in obstacles:
            if (so - s_cur) < BRAKE_LOOKAHEAD:
                brake = True
        uref = tab.get_control(s_cur)
        if brake:
            g.append([uref[0], BRAKE_GUESS])
        else:
            g.append([uref[0], uref[1]])
        s_cur += v_cur * DT
    return np.array(g).ravel()


def bounds_list():
    """trajectory_tracking.py:249."""
    return [(U_MIN[0], U_MAX[0]), (U_MIN[1], U_MAX[1])] * N


def _norm_obs(obstacles):
    out = []
    for o in obstacles:
        if isinstance(o, dict):
            out.append((float(o["s"]), float(o["v"])))
        else:
            out.append((float(o[0]), float(o[1])))
    return out


def solve_as_shipped(tab: RefTable, x0, obstacles):
    """Reference ``solve`` exactly as shipped (trajectory_tracking.py:213-263):
    SLSQP, ftol=1e-3, maxiter=15, finite-difference gradients."""
    obstacles = _norm_obs(obstacles)
    x0 = np.asarray(x0, dtype=np.float64)
    U0 = warm_start(tab, x0, obstacles)
    cons = {"type": "ineq", "fun": lambda U: constraint_values(tab, U, x0, obstacles)}
    t0 = time.time()
    sol = minimize(lambda U, x: cost(tab, U, x), U0, args=(x0,), method="SLSQP",
                   bounds=bounds_list(), constraints=cons,
                   options={"ftol": 1e-3, "disp": False, "maxiter": 15})
    t1 = time.time()
    U = sol.x.reshape(N, 2)
    return U[0], predict(tab, x0, sol.x), t1 - t0, sol


def solve_converged(tab: RefTable, x0, obstacles, U_start=None, jac=None, ftol=1e-12, maxiter=500):
    """Reference formulation solved to convergence."""
    obstacles = _norm_obs(obstacles)
    x0 = np.asarray(x0, dtype=np.float64)
    U0 = warm_start(tab, x0, obstacles) if U_start is None else np.asarray(U_start, dtype=np.float64)
    cons = {"type": "ineq", "fun": lambda U: constraint_values(tab, U, x0, obstacles)}
    kw = {}
    if jac is not None:
        kw["jac"] = jac
    sol = minimize(lambda U, x: cost(tab, U, x), U0, args=(x0,), method="SLSQP",
                   bounds=bounds_list(), constraints=cons,
                   options={"ftol": ftol, "disp": False, "maxiter": maxiter}, **kw)
    return sol


def converged_oracle(tab: RefTable, x0, obstacles, agree_tol=5e-5, feas_tol=1e-8):
    """Acceptance rule of SURVEY.md §8c: run twice (2-point and 3-point differences, the
    second started from the first), accept when both statuses are in {0, 8}, min c >= -feas_tol
    and the two answers agree within ``agree_tol``.  Returns dict."""
    obstacles = _norm_obs(obstacles)
    a = solve_converged(tab, x0, obstacles)
    b = solve_converged(tab, x0, obstacles, U_start=a.x, jac="3-point")
    cb = constraint_values(tab, b.x, x0, obstacles)
    ca = constraint_values(tab, a.x, x0, obstacles)
    agree = float(np.max(np.abs(a.x - b.x)))
    pinned = (a.status in (0, 8) and b.status in (0, 8)
              and ca.min() >= -feas_tol and cb.min() >= -feas_tol and agree <= agree_tol)
    best = b if b.fun <= a.fun else a
    return {"U": best.x.copy(), "J": float(best.fun), "status": (int(a.status), int(b.status)),
            "agree": agree, "min_c": float(min(ca.min(), cb.min())), "pinned": bool(pinned),
            "c": constraint_values(tab, best.x, x0, obstacles)}


# --------------------------------------------------------------------------------------
# Environment: obstacle FSM + closed loop.  trajectory_tracking.py:266-443
# --------------------------------------------------------------------------------------
FSM_TRAJ2 = dict(obs_trigger_s=710.0, obs_start_s=780.0, obs_v=4.0, obs_end_s=1050.0,
                 tl_pos=550.0, tl_trigger_s=100.0, tl_stop_duration=20.0)   # :294-308
FSM_TRAJ3 = dict(obs_trigger_s=5.0, obs_start_s=150.0, obs_v=4.0, obs_end_s=850.0,
                 tl_pos=2000.0, tl_trigger_s=100.0, tl_stop_duration=20.0)  # :313-327 (commented block)


class ObstacleFSMPort:
    """trajectory_tracking.py:285-374, with the scenario constants as arguments."""

    def __init__(self, dynamic_obstacle=False, traffic_light=False, **c):
        cfg = dict(FSM_TRAJ2)
        cfg.update(c)
        self.dynamic_obstacle = dynamic_obstacle
        self.traffic_light = traffic_light
        self.obs_trigger_s = cfg["obs_trigger_s"]
        self.obs_start_s = cfg["obs_start_s"]
        self.obs_v = cfg["obs_v"]
        self.obs_end_s = cfg["obs_end_s"]
        self.obs_active = False
        self.obs_s = self.obs_start_s
        self.obs_has_triggered = False
        self.tl_pos = cfg["tl_pos"]
        self.tl_trigger_s = cfg["tl_trigger_s"]
        self.tl_stop_duration = cfg["tl_stop_duration"]
        self.tl_state = "RED"
        self.tl_timer = 0.0
        self.tl_waiting = False

    def update(self, dt, s, v):
        act = []
        if self.dynamic_obstacle:
            if (not self.obs_has_triggered) and s >= self.obs_trigger_s:
                self.obs_has_triggered = True
                self.obs_active = True
            if self.obs_active:
                self.obs_s += self.obs_v * dt
                if self.obs_s > self.obs_end_s:
                    self.obs_active = False
                else:
                    act.append({"s": self.obs_s, "v": self.obs_v, "type": "car"})
        if self.traffic_light:
            dist = self.tl_pos - s
            if self.tl_state == "RED":
                if 0 < dist < self.tl_trigger_s:
                    act.append({"s": self.tl_pos, "v": 0.0, "type": "light"})
                    if v < 0.1 and dist < 10.0:
                        self.tl_waiting = True
                if self.tl_waiting:
                    self.tl_timer += dt
                    if self.tl_timer >= self.tl_stop_duration:
                        self.tl_state = "GREEN"
                        self.tl_waiting = False
        return act, self.tl_state


def closed_loop(tab: RefTable, fsm, solve_fn, max_steps=100000):
    """run_simulation, trajectory_tracking.py:377-443, minus printing/plots.
    ``solve_fn(x, obstacles) -> (u0, pred_X, seconds)``."""
    x = np.array([0.0, 0.0, 0.0, 0.0, 0.5])
    cur_s = x[0]
    hx, hu, ht, hobs, htl, hn = [x], [], [], [], [], []
    step = 0
    while cur_s <= tab.s_max - 1.0 and step < max_steps:
        obstacles, tl = fsm.update(DT, x[0], x[4])
        u, _pred, sec = solve_fn(x, obstacles)
        k_ref = tab.get_state(cur_s)[3]
        x = x + DT * dynamics(x, u, k_ref)
        cur_s = x[0]
        hx.append(x)
        hu.append(np.array(u, dtype=np.float64))
        ht.append(sec)
        htl.append(tl)
        hobs.append(next((o["s"] for o in obstacles if o["type"] == "car"), np.nan))
        hn.append(len(obstacles))
        step += 1
    return {"x": np.array(hx), "u": np.array(hu), "t": np.array(ht), "obs_s": np.array(hobs),
            "tl": htl, "n_obs": np.array(hn)}


def tracking_verdicts(hist, fsm, s_total):
    """sanity_checks.py:79-184 as a dict of booleans (CPU-time item reported separately)."""
    hx, hu = hist["x"], hist["u"]
    v = {}
    v["destination"] = bool(hx[-1, 0] >= s_total - 1.0)
    v["on_road"] = bool(np.max(np.abs(hx[:, 1])) <= 1.5)
    v["steer_ok"] = not ((hu[:, 0].min() < U_MIN[0] - 0.1) or (hu[:, 0].max() > U_MAX[0] + 0.1))
    v["accel_ok"] = not ((hu[:, 1].min() < U_MIN[1] - 0.1) or (hu[:, 1].max() > U_MAX[1] + 0.1))
    if fsm.dynamic_obstacle:
        obs = hist["obs_s"]
        m = ~np.isnan(obs)
        if m.any():
            L = min(len(hx), len(obs))
            d = obs[:L][m[:L]] - hx[:L, 0][m[:L]]
            v["obstacle_avoided"] = bool(d.min() >= 1.0)
    if fsm.traffic_light:
        idx = np.where(hx[:, 0] > fsm.tl_pos)[0]
        viol = False
        if len(idx) > 0 and idx[0] < len(hist["tl"]) and hist["tl"][idx[0]] == "RED":
            viol = True
        v["light_respected"] = not viol
    return v


# --------------------------------------------------------------------------------------
# Monte-Carlo problem generator -- SURVEY.md §8(d) config 4 (synthetic; not in the reference)
# --------------------------------------------------------------------------------------
MC_SEED = 20261018


def monte_carlo_problems(tab: RefTable, B, seed=MC_SEED):
    """Returns x0[B,5], obs_sv[B,2,2], n_obs[B] (float64 / int32)."""
    rng = np.random.default_rng(seed)
    s0 = rng.uniform(0.0, tab.s_max - 30.0, size=B)
    nd = rng.normal(0.0, 0.03, size=B)
    no = rng.normal(0.0, 0.01, size=B)
    nk = rng.normal(0.0, 0.005, size=B)
    nv = rng.normal(0.0, 0.5, size=B)
    cls = rng.choice(4, size=B, p=[0.50, 0.25, 0.15, 0.10])
    ug1 = rng.uniform(size=B)
    ug2 = rng.uniform(size=B)
    uv = rng.uniform(2.0, 8.0, size=B)
    us = rng.uniform(size=B)
    x0 = np.zeros((B, 5))
    obs = np.zeros((B, 2, 2))
    n_obs = np.zeros(B, dtype=np.int32)
    # vectorised table lookup (same formula as RefTable._lin)
    i = np.clip(np.searchsorted(tab.s, s0, side="left"), 1, tab.K - 1)
    lo = i - 1
    wl = (s0 - tab.s[lo]) / (tab.s[i] - tab.s[lo])
    wr = (tab.s[i] - s0) / (tab.s[i] - tab.s[lo])
    ref = wl[:, None] * tab.X[i] + wr[:, None] * tab.X[lo]
    x0[:, 0] = s0
    x0[:, 1] = np.clip(ref[:, 1], -0.3, 0.3) + nd
    x0[:, 2] = np.clip(ref[:, 2], -0.1, 0.1) + no
    x0[:, 3] = ref[:, 3] + nk
    x0[:, 4] = np.clip(ref[:, 4] + nv, 0.2, 14.0)
    v0 = x0[:, 4]
    stress = (np.arange(B) % 20) == 0
    lo_gap = 1.6 * v0 + 6.0
    gap_car = lo_gap + ug1 * np.maximum(60.0 - lo_gap, 1.0)
    gap_light = lo_gap + ug2 * np.maximum(100.0 - lo_gap, 1.0)
    gap_stress = 2.0 + us * np.maximum(1.5 * v0 - 2.0, 0.1)
    has_car = (cls == 1) | (cls == 3)
    has_light = (cls == 2) | (cls == 3)
    # stress slice: force at least one obstacle, too close to be feasible
    force = stress & ~(has_car | has_light)
    has_car = has_car | force
    gap_car = np.where(stress & has_car, gap_stress, gap_car)
    gap_light = np.where(stress & has_light & ~has_car, gap_stress, gap_light)
    car_s, light_s = s0 + gap_car, s0 + gap_light
    both = has_car & has_light
    obs[:, 0, 0] = np.where(has_car, car_s, np.where(has_light, light_s, 0.0))
    obs[:, 0, 1] = np.where(has_car, uv, 0.0)
    obs[:, 1, 0] = np.where(both, light_s, 0.0)
    n_obs[:] = has_car.astype(np.int32) + has_light.astype(np.int32)
    return x0, obs, n_obs
