"""Development: first-pass / robust-pass times of the thread-per-problem kernels on small batches (a closed-loop step of a
few thousand vehicles, a shard of a strong-scaling run)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz")
T = M.BatchedTracker(L, coop_max_batch=0)
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
x0, obs, n = P.monte_carlo_problems(tab, 65536)
for B in (2048, 4096, 8192, 16384, 32768, 65536):
    d = [torch.from_numpy(a[:B]).cuda() for a in (x0, obs, n)]
    out = T.solve_batch(*d)
    torch.cuda.synchronize()
    ms = []
    for _ in range(6):
        T.solve_batch(*d, out=out); torch.cuda.synchronize(); ms.append(T.last_pass_ms())
    a = np.array([(m[0], m[1]) for m in ms[2:]]).mean(axis=0)
    print(f"B {B}: first pass {a[0]:.3f} ms, robust pass {a[1]:.3f} ms ({ms[-1][2]} left)", flush=True)
