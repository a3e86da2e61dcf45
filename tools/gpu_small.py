"""Development: latency of small batches, thread-per-problem vs warp-per-problem first pass."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz")
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
x0, obs, n = P.monte_carlo_problems(tab, 65536)
for cmb in (0, 1 << 30):
    T = M.BatchedTracker(L, coop_max_batch=cmb)
    for B in (1, 32, 256, 1024, 2048, 3072, 4096, 6144, 8192, 16384):
        dx, do, dn = (torch.from_numpy(a[:B]).cuda() for a in (x0, obs, n))
        out = T.solve_batch(dx, do, dn)
        torch.cuda.synchronize()
        ms = []
        for rep in range(20):
            T.solve_batch(dx, do, dn, out=out)
            torch.cuda.synchronize()
            ms.append(T.last_kernel_ms())
        ref = out["U"].cpu().numpy().copy()
        print(f"coop_max_batch={cmb} B={B}: device ms median {np.median(ms):.4f} min {np.min(ms):.4f} passes {T.last_pass_ms()}")
    # host API B=1 latency
    lat = []
    for i in range(220):
        o = [{"s": float(obs[i, k, 0]), "v": float(obs[i, k, 1]), "type": "car"} for k in range(int(n[i]))]
        t0 = time.perf_counter(); T.solve(x0[i], o); lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.array(lat[20:])
    print(f"   B=1 host API latency p50 {np.median(lat):.4f} p99 {np.quantile(lat, .99):.4f} ms")
