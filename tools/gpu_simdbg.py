import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import safe_autonomous_driving_mpc_b200 as M
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory2.npz"); T = M.BatchedTracker(L)
rng = np.random.default_rng(11)
B = 256
x_init = np.tile([0.0, 0.0, 0.0, 0.0, 0.5], (B, 1))
x_init[8:, 1] += rng.normal(0, 0.05, B - 8)
x_init[8:, 4] += rng.uniform(0, 3.0, B - 8)
scen = []
for b in range(B):
    if b < 8: scen.append(M.make_scenario(2))
    else: scen.append(M.make_scenario(2, obs_v=float(rng.uniform(3.0, 6.0)), tl_pos=float(rng.uniform(450.0, 650.0)), tl_stop_duration=float(rng.uniform(5.0, 25.0))))
sim = M.BatchedSimulation(T, scen, B=B, x_init=x_init, history_steps=3000)
sim.run(max_steps=3000, check_every=100)
x, steps, uns = sim.state(); h = sim.history()
al = np.where(x[:, 0] <= L.s_max - 1)[0]
print("alive", al)
for b in al[:6]:
    print(b, "x", np.round(x[b], 3), "tl_pos", scen[b].tl_pos, "obs_v", scen[b].obs_v, "dur", scen[b].tl_stop_duration, "unsolved", uns[b], "last obs", h["obs_s"][-1, b], "tl", h["tl"][-1, b], "status tail", h["status"][-5:, b], "u", h["u"][-1, b])
    k = np.where(np.abs(np.diff(h["x"][:, b, 0])) < 1e-6)[0]
    print("   first stall step", k[:1], "x there", np.round(h["x"][k[0], b], 3) if len(k) else None)
