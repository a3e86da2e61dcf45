"""Quick GPU sanity script (development): parity numbers against tests/golden + a timing of the 65,536 batch."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P, sqp_admm_model as A

def main():
    for ti in (1, 2, 3):
        L = M.TrajectoryLoader(f"{ROOT}/data/trajectory{ti}.npz")
        T = M.BatchedTracker(L)
        z = np.load(f"{ROOT}/tests/golden/fn_traj{ti}.npz")
        r = T.eval_batch(z["x0"], z["U"], z["obs_sv"], z["n_obs"])
        print(f"traj{ti} fn: predict {np.abs(r['Xpred']-z['predict']).max():.2e} cost rel {np.max(np.abs(r['cost']-z['cost'])/np.abs(z['cost'])):.2e} "
              f"cons {np.nanmax(np.abs(r['cons']-z['constraints'])):.2e} nanmatch {np.array_equal(np.isnan(r['cons']),np.isnan(z['constraints']))}")
        tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory{ti}.npz")
        asm = A.assemble(tab, z["x0"], z["U"]); qp = A.build_qp(asm, z["x0"], z["U"], z["obs_sv"], z["n_obs"])
        Hm = np.array([[qp["P"][b][i, j] for i in range(10) for j in range(i + 1)] for b in range(len(z["x0"]))])
        print(f"   lin: H {np.abs(r['lin'][:, :55]-Hm).max():.2e} (scale {np.abs(Hm).max():.1e}) q {np.abs(r['lin'][:,55:65]-qp['q']).max():.2e} "
              f"D {np.abs(r['lin'][:,65:105].reshape(-1,4,10)-asm['dX'][:,2:6,1]).max():.2e} O {np.abs(r['lin'][:,105:145].reshape(-1,4,10)-asm['dX'][:,2:6,2]).max():.2e}")
        for name in ([f"solve_traj{ti}"] + (["solve_mc_traj3"] if ti == 3 else [])):
            g = np.load(f"{ROOT}/tests/golden/{name}.npz")
            w = T.eval_batch(g["x0"], np.zeros((len(g["x0"]), 10)), g["obs_sv"], g["n_obs"])["warm"]
            t0 = time.time(); s = T.solve_batch_host(g["x0"], g["obs_sv"], g["n_obs"]); dt = time.time() - t0
            pin = g["pinned"]
            err = np.abs(s["U"].reshape(-1, 10) - g["U_conv"]).max(axis=1)
            jr = np.abs(s["obj"] - g["J_conv"]) / np.maximum(np.abs(g["J_conv"]), 1.0)
            print(f"   {name}: warm {np.abs(w-g['U_init']).max():.1e} | pinned {pin.sum()}/{len(pin)} err max {err[pin].max():.2e} n>1e-4 {(err[pin]>1e-4).sum()} Jrel max {jr[pin].max():.1e} "
                  f"| status pinned {np.bincount(s['status'][pin],minlength=3)} unpinned {np.bincount(s['status'][~pin],minlength=3)} | rounds {s['iters'][pin,0].mean():.2f} iters {s['iters'][pin,1].mean():.0f}/{s['iters'][:,1].max()} | {dt*1e3:.1f} ms kernel {T.last_kernel_ms():.2f} ms")
            bad = np.where(pin & (err > 1e-4))[0]
            for b in bad[:5]:
                print("      bad", b, err[b], s["status"][b], s["iters"][b], s["obj"][b], g["J_conv"][b])
    # throughput
    L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz"); T = M.BatchedTracker(L)
    tab = P.RefTable.from_npz(f"{ROOT}/data/trajectory3.npz")
    x0, obs, n = P.monte_carlo_problems(tab, 65536)
    for rep in range(3):
        t0 = time.time(); s = T.solve_batch_host(x0, obs, n); dt = time.time() - t0
        print(f"MC 65536: e2e {dt*1e3:.1f} ms kernel {T.last_kernel_ms():.2f} ms -> {65536/(T.last_kernel_ms()*1e-3):.3e} solves/s; status {np.bincount(s['status'],minlength=3)} rounds {s['iters'][:,0].mean():.2f} iters mean {s['iters'][:,1].mean():.0f} max {s['iters'][:,1].max()} passes {T.last_pass_ms()}")
    tf, ms = T.measure_fp64_peak()
    print(f"fp64 DFMA peak {tf:.2f} TFLOP/s ({ms:.2f} ms)")
    x1 = x0[:1]; 
    ts = []
    for i in range(200):
        t0 = time.perf_counter(); T.solve_batch_host(x0[i:i+1], obs[i:i+1], n[i:i+1]); ts.append(time.perf_counter() - t0)
    ts = np.array(ts[20:]) * 1e3
    print(f"B=1 latency ms: p50 {np.median(ts):.3f} p99 {np.quantile(ts,.99):.3f} max {ts.max():.3f}")

if __name__ == "__main__":
    main()
