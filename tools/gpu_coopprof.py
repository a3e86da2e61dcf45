"""Development: phase split of the cooperative kernel (library built with -DMPCB_COOP_PROFILE, MPCB_LIB set)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P
lib = M._lib.load()
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz"); T = M.BatchedTracker(L)
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
x0, obs, n = P.monte_carlo_problems(tab, 65536)
names = ["prologue", "linearise", "rows", "factor", "iterations", "round_end", "total", "n/iters*1000+rounds"]
def run(B):
    dx, do, dn = (torch.from_numpy(a[:B]).cuda() for a in (x0, obs, n))
    out = T.solve_batch(dx, do, dn); torch.cuda.synchronize()
    s = (C.c_ulonglong * 8)(); m = (C.c_ulonglong * 8)()
    lib.mpcb_debug_coop_profile(s, m, 1)
    T.solve_batch(dx, do, dn, out=out); torch.cuda.synchronize()
    lib.mpcb_debug_coop_profile(s, m, 0)
    print(f"B={B} passes {T.last_pass_ms()}")
    print("  mean per warp (cycles):", {k: int(s[i] / max(s[7], 1)) for i, k in enumerate(names[:7])}, "warps", s[7])
    print("  slowest warp  (cycles):", {k: int(m[i]) for i, k in enumerate(names)})
run(65536); run(1)
