"""Development: small run of every kernel of libmpcb200.so for compute-sanitizer (memcheck / racecheck / initcheck):
the two solve kernels in both execution shapes and both host paths (packed, chunked, asynchronous with three handles),
the evaluation kernel, the planner kernels, and the closed-loop kernels (FSM, plant, alive count, checks).
    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz")
T = M.BatchedTracker(L)
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
SMALL = bool(os.environ.get("MPCB_SAN_SMALL"))     # racecheck: ~100x slower
x0, obs, n = P.monte_carlo_problems(tab, 6000)
for B in ((1, 33, 600, 3200) if SMALL else (1, 33, 2500, 6000)):
    r = T.solve_batch_host(x0[:B], obs[:B], n[:B])
    print(B, np.bincount(r["status"], minlength=3), flush=True)
Tb = M.BatchedTracker(L, coop_max_batch=0, coop_pass2=0)            # thread-per-problem kernels, both passes
r = Tb.solve_batch_host(x0[:700], obs[:700], n[:700])
print("bulk", np.bincount(r["status"], minlength=3), flush=True)
u = T.solve_batch_host_u0(x0[:4000], obs[:4000], n[:4000], want_obj=True)
PB = M.tracker.PinnedBuffer
Ts = [M.BatchedTracker(L) for _ in range(3)]
pins = [[PB(a[:4000].shape, a.dtype) for a in (x0, obs, n)] for _ in range(3)]
outs = [dict(U=PB((4000, 5, 2), np.float64), Xpred=PB((4000, 6, 5), np.float64), status=PB((4000,), np.int32)) for _ in range(3)]
for rnd in range(3):
    for k in range(3):
        for b, a in zip(pins[k], (x0, obs, n)):
            b.array[...] = a[:4000]
        Ts[k].solve_batch_host_async(*[b.array for b in pins[k]], {kk: o.array for kk, o in outs[k].items()})
    for k in range(3):
        Ts[k].wait()
print("async", np.bincount(outs[0]["status"].array, minlength=3), flush=True)
g = np.load(f"{ROOT}/tests/golden/solve_traj2.npz")
r = T.solve_batch_host(g["x0"], g["obs_sv"], g["n_obs"])
e = T.eval_batch(x0[:100], np.zeros((100, 10)), obs[:100], n[:100])
z3 = np.load(f"{ROOT}/data/trajectory3.npz")
E = M.PlannerEvaluator(T, N=len(z3["U"]), simpson_sign=+1)
z = E.pack(z3["X"], z3["U"], z3["S"])
o = E.evaluate_host(z, lam=np.ones((1, len(z3["U"]), 5)), want_jac=True, want_hess=True)
print("planner", float(np.abs(o["defect"]).max()), flush=True)
# closed loop on the device: per-vehicle scenarios, histories, checks
L2 = M.TrajectoryLoader(f"{ROOT}/data/trajectory2.npz")
T2 = M.BatchedTracker(L2)
rng = np.random.default_rng(3)
scen = [M.make_scenario(2, tl_pos=float(rng.uniform(400, 700)), obs_v=float(rng.uniform(3, 6))) for _ in range(48)]
sim = M.BatchedSimulation(T2, scen, history_steps=400)
sim.step(120 if SMALL else 400)
c = sim.check()
x, steps, uns = sim.state()
h = sim.history()
print("sim", int(sim.alive()), int(steps.max()), int(c["on_road"].sum()), flush=True)
print("ok")
