"""Development: wall time vs device span of the host-buffer entry point on the 65,536 Monte-Carlo batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz")
T = M.BatchedTracker(L)
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
for B in (65536, 16384, 4096):
    x0, obs, n = P.monte_carlo_problems(tab, B)
    pin = {k: M.tracker.PinnedBuffer(a.shape, a.dtype) for k, a in (("x0", x0), ("obs", obs), ("n", n))}
    pin["x0"].array[...] = x0; pin["obs"].array[...] = obs; pin["n"].array[...] = n
    wall, span = [], []
    for i in range(12):
        c0 = T.launch_count()
        t0 = time.perf_counter()
        r = T.solve_batch_host(pin["x0"].array, pin["obs"].array, pin["n"].array)
        wall.append((time.perf_counter() - t0) * 1e3)
        span.append(T.last_kernel_ms())
        if i in (0, 1, 2, 3):
            print(f"  call {i}: wall {wall[-1]:.3f} ms, device span {span[-1]:.3f} ms, launches {T.launch_count() - c0}")
    print(f"B={B}: wall median {np.median(wall[4:]):.3f} ms, device span median {np.median(span[4:]):.3f} ms -> {B / np.median(wall[4:]) / 1e3:.1f} M solves/s")
