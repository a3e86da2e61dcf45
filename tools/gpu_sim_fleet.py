"""Development: device closed loop of a large fleet with staggered starts (vehicle-steps/s as a function of fleet size)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import safe_autonomous_driving_mpc_b200 as M
kw = {}
HOT = False
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "hot":
        HOT = bool(int(v))
    else:
        kw[k] = int(v)
print(kw)
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz"); T = M.BatchedTracker(L, **kw)
rng = np.random.default_rng(11)
for B in (4096, 65536):
    xi = np.zeros((B, 5))
    s0 = rng.uniform(0.0, L.s_max - 400.0, B)
    for b in range(B):
        xi[b] = L.get_state(s0[b])
    xi[:, 4] = np.clip(xi[:, 4], 0.5, None)
    scen = [M.make_scenario(3, tl_pos=float(s + rng.uniform(150, 350)), obs_trigger_s=float(s + 5), obs_start_s=float(s + 60),
                            obs_end_s=float(s + 300)) for s in s0]
    sim = M.BatchedSimulation(T, scen, x_init=xi, hot_start=HOT)
    sim.step(8); torch.cuda.synchronize()
    x0_, st0, _ = sim.state()
    t0 = time.perf_counter()
    sim.step(400)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    x, steps, uns = sim.state()
    print("   last step: first pass %.3f ms, robust pass %.3f ms, %d vehicles left to it" % T.last_pass_ms(), flush=True)
    print(f"fleet {B}: 400 steps in {dt * 1e3:.1f} ms -> {(steps.sum() - st0.sum()) / dt / 1e6:.1f} M vehicle-steps/s, alive {sim.alive()}, unsolved steps {int(uns.sum())}", flush=True)
    c = None
