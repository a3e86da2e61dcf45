"""Generate golden vectors from the UNMODIFIED reference (imported from /root/reference).

Build-container only.  Writes tests/golden/{fn,solve,closed_loop}_traj{1,2,3}.npz and
tests/golden/solve_mc_traj3.npz.  Every array in those files is produced by the reference's own
TrajectoryLoader / TrajectoryTracker / ObstaclesFSM objects (and scipy's SLSQP called on the
reference's own cost/constraints callables for the converged answers).  The versions of numpy /
scipy used are stored in each file.

    python tools/make_golden.py [--procs 8] [--only fn|solve|closed_loop|mc]
"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np
import scipy

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

VERS = np.array([np.__version__, scipy.__version__])

FSM_CFG = {
    1: dict(dynamic_obstacle=False, traffic_light=False, const=None),
    2: dict(dynamic_obstacle=True, traffic_light=True, const=None),  # constants as committed (:294-308)
    3: dict(dynamic_obstacle=True, traffic_light=True,                # commented block (:313-327)
            const=dict(obs_trigger_s=5.0, obs_start_s=150.0, obs_v=4.0, obs_end_s=850.0, obs_s=150.0,
                       tl_pos=2000.0, tl_trigger_s=100.0, tl_stop_duration=20.0)),
}

_T = {}


def _tracker(i):
    if i not in _T:
        import trajectory_tracking as tt
        from trajectory_loader import TrajectoryLoader
        L = TrajectoryLoader(os.path.join(REF, "trajectories", f"trajectory{i}.json"))
        _T[i] = (L, tt.TrajectoryTracker(L), tt)
    return _T[i]


def _obs_list(obs_sv, n):
    return [{"s": float(obs_sv[k, 0]), "v": float(obs_sv[k, 1]), "type": "car"} for k in range(int(n))]


# ----------------------------------------------------------------------------- function vectors
def make_fn(i, n=96):
    L, T, tt = _tracker(i)
    rng = np.random.default_rng(1000 + i)
    s_max = L.s_max
    sq = np.concatenate([rng.uniform(-10.0, s_max + 10.0, 200),
                         [0.0, -1.0, s_max, s_max - 1e-9, s_max + 1.0],
                         L.X_ref[[0, 1, 5, -3, -2, -1], 0],          # exact knot hits
                         L.X_ref[-2, 0] + np.array([1e-3, 0.05])])   # past last control knot
    gs = np.array([L.get_state(s) for s in sq])
    gc = np.array([L.get_control(s) for s in sq])
    x0 = np.zeros((n, 5)); U = np.zeros((n, 10)); obs = np.zeros((n, 2, 2)); n_obs = np.zeros(n, np.int32)
    pred = np.zeros((n, 6, 5)); cost = np.zeros(n); cons = np.full((n, 45), np.nan)
    for t in range(n):
        s0 = rng.uniform(0.0, s_max - 2.0) if t % 8 else rng.uniform(s_max - 12.0, s_max + 1.0)
        x0[t] = [s0, rng.normal(0, 0.15), rng.normal(0, 0.06), rng.normal(0, 0.05), rng.uniform(-0.5, 14.0)]
        U[t] = rng.normal(0, 1, 10) * np.tile([0.3, 2.5], 5)
        n_obs[t] = rng.integers(0, 3)
        for k in range(n_obs[t]):
            obs[t, k] = [s0 + rng.uniform(-5.0, 60.0), rng.uniform(0.0, 8.0)]
        ol = _obs_list(obs[t], n_obs[t])
        pred[t] = T.predict(x0[t], U[t])
        cost[t] = T.cost(U[t], x0[t])
        c = T.constraints(x0[t], ol)["fun"](U[t])
        cons[t, : len(c)] = c
    np.savez_compressed(os.path.join(OUT, f"fn_traj{i}.npz"), versions=VERS, s_query=sq, get_state=gs,
                        get_control=gc, x0=x0, U=U, obs_sv=obs, n_obs=n_obs, predict=pred, cost=cost,
                        constraints=cons)
    print("fn", i, "done", flush=True)


# ----------------------------------------------------------------------------- solve vectors
def _solve_one(args):
    """as-shipped solve + converged oracle (2-point then 3-point from the 2-point answer)."""
    i, x0, obs_sv, n = args
    L, T, tt = _tracker(i)
    from scipy.optimize import minimize
    ol = _obs_list(obs_sv, n)
    rec = {}
    real_min = tt.minimize

    def spy(fun, U_init, **kw):
        rec["U_init"] = np.array(U_init, dtype=float).copy()
        r = real_min(fun, U_init, **kw)
        rec["res"] = r
        return r

    tt.minimize = spy
    try:
        u0, predX, _sec = T.solve(x0, ol)
    finally:
        tt.minimize = real_min
    r = rec["res"]
    bounds = [(T.u_min[0], T.u_max[0]), (T.u_min[1], T.u_max[1])] * T.N
    cons = T.constraints(x0, ol)
    opt = {"ftol": 1e-12, "disp": False, "maxiter": 500}
    a = minimize(T.cost, rec["U_init"], args=(x0,), method="SLSQP", bounds=bounds, constraints=cons, options=opt)
    b = minimize(T.cost, a.x, args=(x0,), method="SLSQP", jac="3-point", bounds=bounds, constraints=cons,
                 options=opt)
    ca = cons["fun"](a.x); cb = cons["fun"](b.x)
    best = b if b.fun <= a.fun else a
    cbest = cons["fun"](best.x)
    cpad = np.full(45, np.nan); cpad[: len(cbest)] = cbest
    return dict(U_init=rec["U_init"], U_ship=r.x.copy(), ship_status=int(r.status), ship_nit=int(r.nit),
                ship_fun=float(r.fun), predX_ship=predX,
                U_conv=best.x.copy(), J_conv=float(best.fun), st_a=int(a.status), st_b=int(b.status),
                agree=float(np.max(np.abs(a.x - b.x))), min_c=float(min(ca.min(), cb.min())), c_conv=cpad)


def _pack_solve(path, i, x0, obs, n_obs, results, extra=None):
    keys = results[0].keys()
    d = {k: np.array([r[k] for r in results]) for k in keys}
    pinned = (np.isin(d["st_a"], (0, 8)) & np.isin(d["st_b"], (0, 8)) & (d["min_c"] >= -1e-8)
              & (d["agree"] <= 5e-5))
    d.update(versions=VERS, traj=np.int32(i), x0=x0, obs_sv=obs, n_obs=n_obs, pinned=pinned)
    if extra:
        d.update(extra)
    np.savez_compressed(path, **d)
    print(path, "n", len(results), "pinned", int(pinned.sum()), flush=True)


def make_closed_loop(i):
    """Reference run_simulation (unmodified) with per-step obstacle sets and SLSQP status recorded."""
    L, T, tt = _tracker(i)
    cfg = FSM_CFG[i]
    fsm = tt.ObstaclesFSM(dynamic_obstacle=cfg["dynamic_obstacle"], traffic_light=cfg["traffic_light"])
    if cfg["const"]:
        for k, v in cfg["const"].items():
            setattr(fsm, k, v)
    obs_log, st_log = [], []
    real_update = fsm.update
    real_min = tt.minimize

    def upd(dt, s, v):
        o, tl = real_update(dt, s, v)
        obs_log.append([(d["s"], d["v"]) for d in o])
        return o, tl

    def spy(fun, U_init, **kw):
        r = real_min(fun, U_init, **kw)
        st_log.append((int(r.status), int(r.nit), int(r.nfev)))
        return r

    fsm.update = upd
    tt.minimize = spy
    import io, contextlib
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            hx, hu, ht, hp, hobs, htl, _ = tt.run_simulation(T, fsm, L)
    finally:
        tt.minimize = real_min
    n = len(hu)
    obs = np.zeros((n, 2, 2)); n_obs = np.zeros(n, np.int32)
    for t, o in enumerate(obs_log):
        n_obs[t] = len(o)
        for k, (s, v) in enumerate(o):
            obs[t, k] = (s, v)
    np.savez_compressed(os.path.join(OUT, f"closed_loop_traj{i}.npz"), versions=VERS, hist_x=hx, hist_u=hu,
                        hist_t=ht, hist_pred=np.array(hp), hist_obs_s=np.array(hobs, dtype=float),
                        hist_tl=np.array(htl), obs_sv=obs, n_obs=n_obs, slsqp=np.array(st_log),
                        sanity_stdout=np.array(buf.getvalue()))
    print("closed_loop", i, "steps", n, flush=True)
    return hx, obs, n_obs, np.array(st_log)


def make_solve_from_closed_loop(i, pool, per=72):
    z = np.load(os.path.join(OUT, f"closed_loop_traj{i}.npz"))
    hx, obs, n_obs, st = z["hist_x"], z["obs_sv"], z["n_obs"], z["slsqp"]
    n = len(n_obs)
    rng = np.random.default_rng(2000 + i)
    idx = set(rng.choice(n, size=min(per // 2, n), replace=False).tolist())
    act = np.where(n_obs > 0)[0]
    if len(act):
        idx |= set(rng.choice(act, size=min(per // 2, len(act)), replace=False).tolist())
    bad = np.where(st[:, 0] != 0)[0]
    idx |= set(bad[:12].tolist())
    idx = np.array(sorted(idx))
    res = pool.map(_solve_one, [(i, hx[t].copy(), obs[t].copy(), int(n_obs[t])) for t in idx], chunksize=1)
    _pack_solve(os.path.join(OUT, f"solve_traj{i}.npz"), i, hx[idx], obs[idx], n_obs[idx], res,
                extra=dict(step_index=idx))


def make_solve_mc(pool, n=2048):
    from oracle import tracker_port as P
    tab = P.RefTable.from_npz(os.path.join(ROOT, "data", "trajectory3.npz"))
    x0, obs, n_obs = P.monte_carlo_problems(tab, 65536)
    x0, obs, n_obs = x0[:n], obs[:n], n_obs[:n]
    res = pool.map(_solve_one, [(3, x0[t].copy(), obs[t].copy(), int(n_obs[t])) for t in range(n)], chunksize=4)
    _pack_solve(os.path.join(OUT, "solve_mc_traj3.npz"), 3, x0, obs, n_obs, res)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--only", default="all")
    ap.add_argument("--mc-n", type=int, default=2048, help="Monte-Carlo problems with a converged answer (BASELINE.md 3)")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    with mp.Pool(a.procs) as pool:
        if a.only in ("all", "fn"):
            pool.map(make_fn, [1, 2, 3])
        if a.only in ("all", "closed_loop"):
            pool.map(make_closed_loop, [3, 2, 1])
        if a.only in ("all", "mc"):
            make_solve_mc(pool, a.mc_n)
        if a.only in ("all", "solve"):
            for i in (1, 2, 3):
                make_solve_from_closed_loop(i, pool)
