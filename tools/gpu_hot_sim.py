import os, sys, time
ROOT = "/root/repo" if os.path.exists("/root/repo/data") else os.getcwd()
sys.path.insert(0, ROOT)
import numpy as np, torch
import safe_autonomous_driving_mpc_b200 as M
# synchronized fleets (the bench config): step counts and verdicts with and without the hot start
for i in (1, 2, 3):
    L = M.TrajectoryLoader(f"{ROOT}/data/trajectory{i}.npz"); T = M.BatchedTracker(L)
    scen = M.make_scenario(2, dynamic_obstacle=0, traffic_light=0) if i == 1 else M.make_scenario(i)
    for B in (1, 4096):
        res = []
        for hot in (False, True):
            sim = M.BatchedSimulation(T, scen, B=B, history_steps=2500 if B == 1 else 0, hot_start=hot)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            sim.run(max_steps=2600, check_every=128)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            x, steps, uns = sim.state()
            c = sim.check() if B == 1 else None
            res.append((steps.copy(), uns.copy(), x.copy()))
            print(f"traj{i} B={B} hot={hot}: steps {steps[0]} unsolved {uns[0]} {steps.sum() / dt / 1e6:.1f} M vehicle-steps/s" + (f" passed {bool(c['passed'][0])}" if c else ""), flush=True)
        print("    |dx final| max", np.abs(res[0][2] - res[1][2]).max(), "step diff", np.abs(res[0][0] - res[1][0]).max())
