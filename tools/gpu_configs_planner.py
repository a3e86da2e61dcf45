"""Development: timing of the planner Hermite-Simpson evaluator only (BASELINE config 5): trajectory3 as one batch of
1,258 intervals and tiled x64 (80,512), defects + Jacobian + Hessian blocks; and the Jacobian-only form."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for tiles in (1, 64):
    zt = torch.from_numpy(np.tile(z, (tiles, 1))).cuda()
    lam = torch.from_numpy(np.random.default_rng(7).normal(size=(tiles, N, 5))).cuda()
    for hess in (True, False):
        o = Ev.eval_defects(zt, lam=lam if hess else None, want_jac=True, want_hess=hess)
        ms = []
        for rep in range(14):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); Ev.eval_defects(zt, lam=lam if hess else None, want_jac=True, want_hess=hess, out=o); e1.record()
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        ms = float(np.median(ms[2:]))
        n_int = tiles * N
        byt = n_int * (5 + 60 + (144 + 5 if hess else 0)) * 8 + tiles * (8 * N + 5) * 8
        print(f"x{tiles} hess={hess}: {ms * 1e3:.1f} us  {n_int / ms / 1e6:.3f} G intervals/s  {byt / ms / 1e6:.0f} GB/s = {byt / ms / 1e6 / 6458.7:.3f} of HBM peak")
