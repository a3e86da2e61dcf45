"""Development: a short run of the planner evaluator (x64 tile) and of the device closed loop (4,096 vehicles on
trajectory3, at the red-light approach where a fifth of the solves go to the robust pass) for ncu captures of
mpcb_hs_eval_kernel, mpcb_hs_nodes_kernel, mpcb_fsm_kernel and mpcb_plant_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import safe_autonomous_driving_mpc_b200 as M
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz"); T = M.BatchedTracker(L)
z3 = np.load(f"{ROOT}/data/trajectory3.npz")
N = len(z3["U"])
Ev = M.PlannerEvaluator(T, N=N, simpson_sign=+1)
z = Ev.pack(z3["X"], z3["U"], z3["S"])
zt = torch.from_numpy(np.tile(z, (64, 1))).cuda()
lam = torch.from_numpy(np.random.default_rng(7).normal(size=(64, N, 5))).cuda()
o = Ev.eval_defects(zt, lam=lam, want_jac=True, want_hess=True)
for _ in range(3):
    Ev.eval_defects(zt, lam=lam, want_jac=True, want_hess=True, out=o)
torch.cuda.synchronize()
sim = M.BatchedSimulation(T, M.make_scenario(3), B=4096)
sim.step(24)
torch.cuda.synchronize()
print("ok")
