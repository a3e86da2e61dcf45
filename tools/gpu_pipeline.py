"""Development (measured r02: 1 handle 0.59 ms/batch, 2: 0.446, 3: 0.415, 4: 0.416; running the robust pass on a
high-priority side stream was slower -- 0.47 -- its persistent CTAs take slots from the first pass): throughput of back-to-back 65,536-problem batches with 1..3 handles on as many streams (the robust
pass of one batch overlapping the first pass of the next), rotating over input sets larger than L2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P
dev = torch.device("cuda", 0)
KW = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    KW[k] = int(v)
print(KW)
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz")
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
B, NSET, K = 65536, 8, 40
sets = []
for k in range(NSET):
    x0, obs, n = P.monte_carlo_problems(tab, B, seed=P.MC_SEED + k)
    sets.append([torch.from_numpy(a).to(dev) for a in (x0, obs, n)])
for side in (0,):
    for ns in (1, 2):
        Ts = [M.BatchedTracker(L, **KW) for _ in range(ns)]
        streams = [torch.cuda.Stream(dev) for _ in range(ns)]
        outs = [Ts[0].solve_batch(*sets[k]) for k in range(NSET)]
        torch.cuda.synchronize()
        def run(nsteps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main = torch.cuda.current_stream(dev)
            e0.record(main)
            for s in streams:
                s.wait_event(e0)
            for i in range(nsteps):
                k = i % NSET
                Ts[i % ns].solve_batch(*sets[k], out=outs[k], stream=streams[i % ns].cuda_stream)
            for s in streams:
                ev = torch.cuda.Event(); ev.record(s); main.wait_event(ev)
            e1.record(main)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / nsteps
        run(8)
        ms = [run(K) for _ in range(3)]
        print("   last call per handle (first pass ms, robust pass ms, leftovers):", [tuple(round(v, 3) for v in T.last_pass_ms()) for T in Ts], flush=True)
        print(f"handles {ns}: {min(ms):.4f} ms/batch  {B / min(ms) / 1e3:.1f} M solves/s   (runs {['%.4f' % m for m in ms]})", flush=True)
        del Ts
