"""Development: end-to-end throughput (pinned host buffers in and out) of the asynchronous host entry point with 1..4
handles in flight, full outputs and closed-loop form (U*[0] + status)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import safe_autonomous_driving_mpc_b200 as M
from safe_autonomous_driving_mpc_b200.tracker import PinnedBuffer
from oracle import tracker_port as P
L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz")
tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
B, K = 65536, 60
FULL = dict(U=((B, 5, 2), np.float64), Xpred=((B, 6, 5), np.float64), obj=((B,), np.float64), status=((B,), np.int32),
            iters=((B, 2), np.int32), cmin=((B,), np.float64), active=((B,), np.uint64))
U0 = dict(u0=((B, 2), np.float64), status=((B,), np.int32))
keep = []
def pinned(a):
    b = PinnedBuffer(a.shape, a.dtype); b.array[...] = a; keep.append(b); return b.array
def pinned_out(spec):
    o = {}
    for k, (shp, dt) in spec.items():
        b = PinnedBuffer(shp, dt); keep.append(b); o[k] = b.array
    return o
for name, spec in [x for x in (("full", FULL), ("u0", U0)) if len(sys.argv) < 2 or x[0] in sys.argv[1:]]:
    for nh in (1, 2, 3, 4):
        Ts = [M.BatchedTracker(L) for _ in range(nh)]
        ins = [[pinned(a) for a in P.monte_carlo_problems(tab, B, seed=P.MC_SEED + k)] for k in range(nh)]
        outs = [pinned_out(spec) for _ in range(nh)]
        def run(n):
            t0 = time.perf_counter()
            for i in range(n):
                k = i % nh
                if i >= nh:
                    Ts[k].wait()
                Ts[k].solve_batch_host_async(*ins[k], outs[k])
            for T in Ts:
                T.wait()
            return (time.perf_counter() - t0) / n * 1e3
        run(3 * nh)
        ms = [run(K) for _ in range(6)]
        print(f"{name} handles {nh}: {min(ms):.4f} ms/batch {B / min(ms) / 1e3:.1f} M solves/s median {sorted(ms)[len(ms)//2]:.3f} ({['%.3f' % m for m in ms]}) status {np.bincount(outs[0]['status'], minlength=3).tolist()}", flush=True)
        del Ts
