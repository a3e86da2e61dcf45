"""Development: soak run of the device closed loop -- thousands of vehicles with perturbed starts and per-vehicle
scenario constants on trajectory2 and trajectory3 -- checked with the device sanity check (mpcb_sim_check)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import safe_autonomous_driving_mpc_b200 as M

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kw = {}
for a in sys.argv[2:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v or "e" in v else int(v)
for traj, which, tl_lo, tl_hi in ((2, 2, 450.0, 650.0), (3, 3, 300.0, 900.0)):
    L = M.TrajectoryLoader(f"{ROOT}/data/trajectory{traj}.npz")
    T = M.BatchedTracker(L, **kw)
    rng = np.random.default_rng(100 + traj)
    x_init = np.tile([0.0, 0.0, 0.0, 0.0, 0.5], (B, 1))
    x_init[:, 1] += rng.normal(0, 0.05, B)
    x_init[:, 4] += rng.uniform(0, 3.0, B)
    base = M.make_scenario(which)
    scen = [M.make_scenario(which, obs_v=float(rng.uniform(3.0, 6.0)), tl_pos=float(rng.uniform(tl_lo, tl_hi)),
                            tl_stop_duration=float(rng.uniform(5.0, 25.0))) for _ in range(B)]
    sim = M.BatchedSimulation(T, scen, B=B, x_init=x_init, history_steps=4000, hot_start=bool(int(os.environ.get("HOT", "0"))))
    t0 = time.perf_counter()
    sim.run(max_steps=4000, check_every=200)
    dt = time.perf_counter() - t0
    x, steps, unsolved = sim.state()
    c = sim.check()
    keys = [k for k in c if isinstance(c[k], np.ndarray)]
    print(f"trajectory{traj}: {B} vehicles, alive {sim.alive()}, steps mean {steps.mean():.0f} max {steps.max()}, "
          f"{steps.sum() / dt / 1e6:.1f} M vehicle-steps/s, finite {bool(np.all(np.isfinite(x)))}, "
          f"arrived {(x[:, 0] > L.s_max - 1.0).mean():.4f}, unsolved share {(unsolved / np.maximum(steps, 1)).mean():.4f}")
    stuck = np.nonzero(x[:, 0] <= L.s_max - 1.0)[0]
    for b in stuck[:6]:
        print(f"   stuck vehicle {b}: x = {np.round(x[b], 4).tolist()}, tl_pos {scen[b].tl_pos:.2f}, stop {scen[b].tl_stop_duration:.1f} s, "
              f"obs_v {scen[b].obs_v:.2f}, unsolved {unsolved[b]} of {steps[b]}")
    print("   check:", {k: (float(np.mean(c[k])) if c[k].dtype != np.float64 else float(np.nanmin(c[k]))) for k in keys})
