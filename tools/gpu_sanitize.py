"""Development: small run of every kernel for compute-sanitizer (memcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz")
T = M.BatchedTracker(L)
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
x0, obs, n = P.monte_carlo_problems(tab, 6000)
for B in (1, 33, 2500, 6000):
    r = T.solve_batch_host(x0[:B], obs[:B], n[:B])
    print(B, np.bincount(r["status"], minlength=3))
g = np.load(f"{ROOT}/tests/golden/solve_traj2.npz")
r = T.solve_batch_host(g["x0"], g["obs_sv"], g["n_obs"])
e = T.eval_batch(x0[:100], np.zeros((100, 10)), obs[:100], n[:100])
z3 = np.load(f"{ROOT}/data/trajectory3.npz")
E = M.PlannerEvaluator(T, N=len(z3["U"]), simpson_sign=+1)
z = E.pack(z3["X"], z3["U"], z3["S"])
o = E.evaluate_host(z, lam=np.ones((1, len(z3["U"]), 5)), want_jac=True, want_hess=True)
print("ok", float(np.abs(o["defect"]).max()))
