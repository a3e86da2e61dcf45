"""Development: timing of the 65,536 Monte-Carlo batch only (device-resident), per pass."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P

kw = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    kw[k] = float(v) if "." in v or "e" in v else int(v)
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz")
T = M.BatchedTracker(L, **kw)
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
x0, obs, n = P.monte_carlo_problems(tab, 65536)
dx, do, dn = (torch.from_numpy(a).cuda() for a in (x0, obs, n))
out = T.solve_batch(dx, do, dn)
torch.cuda.synchronize()
for rep in range(4):
    T.solve_batch(dx, do, dn, out=out)
    torch.cuda.synchronize()
    ms = T.last_kernel_ms()
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    print(f"{kw} total {ms:.3f} ms -> {65536 / ms / 1e3:.2f} M solves/s | passes {T.last_pass_ms()} | status "
          f"{np.bincount(st, minlength=3)} rounds {it[:, 0].mean():.2f} iters {it[:, 1].mean():.1f} max {it[:, 1].max()}")
