"""Development: sweep the thread-kernel caps of the first pass (runtime parameters) on the 65,536 Monte-Carlo batch."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P

L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz")
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
x0, obs, n = P.monte_carlo_problems(tab, 65536)
dx, do, dn = (torch.from_numpy(a).cuda() for a in (x0, obs, n))
ref = None
SWEEP = os.environ.get("SWEEP", "thread")
grid = (itertools.product((5, 6), (1, 2, 3), (3, 5)) if SWEEP == "thread" else
        itertools.product((1e4, 1e5, 1e6, 1e7, 1e8), (0,), (1e-9, 1e-6, 1e-4)) if SWEEP == "rho" else
        itertools.product((4, 5, 6, 8, 10, 15), (0,), (1.5, 1.6, 1.7)))
for rounds, segs, its in grid:
    if SWEEP == "thread":
        T = M.BatchedTracker(L, thread_max_rounds=rounds, thread_max_segments=1, fast_segment_iters=segs, thread_fail_rounds=its)
    elif SWEEP == "rho":
        T = M.BatchedTracker(L, fast_rho_on=rounds, fast_rho_off=its)
    else:
        T = M.BatchedTracker(L, segment_iters=rounds, max_segments=max(1, 120 // rounds), alpha=its)
    out = T.solve_batch(dx, do, dn)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(8):
        T.solve_batch(dx, do, dn, out=out)
        torch.cuda.synchronize()
        if T.last_kernel_ms() < best:
            best, passes = T.last_kernel_ms(), T.last_pass_ms()
    st = out["status"].cpu().numpy(); U = out["U"].cpu().numpy()
    if ref is None and (rounds, segs, its) == (3, 1, 2):
        pass
    if (rounds, segs, its) == (6, 4, 2):
        ref = (st.copy(), U.copy())
    print(f"rounds {rounds} segs {segs} fails {its}: {best:.3f} ms  passes {passes[0]:.3f} + {passes[1]:.3f} ({passes[2]})  status {np.bincount(st, minlength=3)}", flush=True)
    del T
