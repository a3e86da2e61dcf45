"""Development script: closed loops on the GPU vs the reference's golden logs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
np.set_printoptions(precision=4, suppress=True, linewidth=200)
import safe_autonomous_driving_mpc_b200 as M
from safe_autonomous_driving_mpc_b200 import environment as E
for i in (1, 2, 3):
    z = np.load(f"{ROOT}/tests/golden/closed_loop_traj{i}.npz")
    L = M.TrajectoryLoader(f"{ROOT}/data/trajectory{i}.npz"); T = M.BatchedTracker(L)
    sc = {1: None, 2: E.SCENARIO_TRAJECTORY2, 3: E.SCENARIO_TRAJECTORY3}[i]
    fsm = M.ObstaclesFSM(i > 1, i > 1, scenario=sc)
    flags = []
    hx, hu, ht, hp, hobs, htl, _ = M.run_simulation(T, fsm, L, record_flags=flags)
    st = np.array([f[0] for f in flags]); n = min(len(hu), len(z["hist_u"]))
    print(f"traj{i}: steps {len(hu)} (ref {len(z['hist_u'])}) status hist {np.bincount(st, minlength=3)} ref fails {np.sum(z['slsqp'][:,0]!=0)}")
    print("   u range", hu.min(axis=0), hu.max(axis=0), "ref", z["hist_u"].min(axis=0), z["hist_u"].max(axis=0))
    print("   v min", hx[:, 4].min(), "ref", z["hist_x"][:, 4].min(), " |d| max", np.abs(hx[:, 1]).max(), "ref", np.abs(z["hist_x"][:, 1]).max())
    print("   non-solved steps", np.where(st != 0)[0][:60])
    print("   ref failed steps", np.where(z["slsqp"][:, 0] != 0)[0][:60])
    d = np.abs(hx[:n] - z["hist_x"][:n]); du = np.abs(hu[:n] - z["hist_u"][:n]).max(axis=1)
    print("   max state diff", d.max(axis=0), "at", d.argmax(axis=0), " du p50/p99/max", np.quantile(du, [.5, .99, 1.0]))
    print(f"   solve ms p50 {np.median(ht)*1e3:.3f} p99 {np.quantile(ht,.99)*1e3:.3f} max {ht.max()*1e3:.3f}")
    bad = np.where(st != 0)[0]
    for t in bad[:25]:
        print("    step", t, "x", hx[t], "u", hu[t], "st", st[t], "ref u", z["hist_u"][t] if t < len(z["hist_u"]) else None)
    if fsm.traffic_light:
        idx = np.where(hx[:, 0] > fsm.tl_pos)[0]
        print("   light passed at step", idx[0] if len(idx) else None, "state then", htl[idx[0]] if len(idx) else None, " min gap to light while red:", min((fsm.tl_pos - hx[t, 0]) for t in range(len(htl)) if htl[t] == "RED"))
    if fsm.dynamic_obstacle:
        o = np.array(hobs, float); m = ~np.isnan(o); print("   min gap to car", (o[m] - hx[:-1, 0][m]).min())
    # dump the per-step problems for offline analysis
    fsm2 = M.ObstaclesFSM(i > 1, i > 1, scenario=sc)
    X0, OBS, NO = [], [], []
    for t in range(len(hu)):
        o, _ = fsm2.update(0.2, hx[t, 0], hx[t, 4])
        ob = np.zeros((2, 2))
        for k, d_ in enumerate(o): ob[k] = (d_["s"], d_["v"])
        X0.append(hx[t]); OBS.append(ob); NO.append(len(o))
    r = T.solve_batch_host(np.array(X0), np.array(OBS), np.array(NO, dtype=np.int32))
    os.makedirs(f"{ROOT}/gpurun_out", exist_ok=True)
    np.savez_compressed(f"{ROOT}/gpurun_out/cl_traj{i}.npz", x0=np.array(X0), obs_sv=np.array(OBS), n_obs=np.array(NO, dtype=np.int32),
                        status=r["status"].copy(), iters=r["iters"].copy(), U=r["U"].copy(), hu=hu, st_loop=st)
    print("   batch re-solve status", np.bincount(r["status"], minlength=3), "iters mean", r["iters"][:, 1].mean(), "max", r["iters"][:, 1].max())
