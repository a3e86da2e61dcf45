"""Measurements for BASELINE.json configs 1-3 and 5 (config 4 is bench.py):
   1-3  closed loops on trajectory1-3 through the reference-shaped call (B = 1 per step): steps, verdicts, per-step latency
        + the same drives as a device loop for 4,096 vehicles at once (vehicle-steps/s)
   5    planner Hermite-Simpson evaluation on trajectory3 (1,258 intervals) and tiled x64 (80,512): intervals/s, GB/s."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import safe_autonomous_driving_mpc_b200 as M
from safe_autonomous_driving_mpc_b200 import environment as E

out = {}
for i in (1, 2, 3):
    L = M.TrajectoryLoader(f"{ROOT}/data/trajectory{i}.npz"); T = M.BatchedTracker(L)
    sc = {1: None, 2: E.SCENARIO_TRAJECTORY2, 3: E.SCENARIO_TRAJECTORY3}[i]
    fsm = M.ObstaclesFSM(i > 1, i > 1, scenario=sc)
    flags = []
    t0 = time.perf_counter()
    hx, hu, ht, hp, hobs, htl, _ = M.run_simulation(T, fsm, L, record_flags=flags)
    wall = time.perf_counter() - t0
    st = np.array([f[0] for f in flags])
    ht = np.array(ht[5:]) * 1e3
    scen = M.make_scenario(2, dynamic_obstacle=0, traffic_light=0) if i == 1 else M.make_scenario(i)
    B = 4096
    sim = M.BatchedSimulation(T, scen, B=B, history_steps=0)
    sim.step(8); torch.cuda.synchronize()
    sim = M.BatchedSimulation(T, scen, B=B, history_steps=0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = sim.run(max_steps=len(hu) + 64, check_every=128)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    x, steps, uns = sim.state()
    out[f"config{i}_trajectory{i}"] = dict(
        steps=len(hu), status_hist=np.bincount(st, minlength=3).tolist(), wall_s=wall,
        solve_ms=dict(p50=float(np.median(ht)), p99=float(np.quantile(ht, .99)), max=float(ht.max())),
        device_loop=dict(vehicles=B, steps_enqueued=int(n), steps_per_vehicle=int(steps[0]), seconds=dt,
                         vehicle_steps_per_s=float(steps.sum() / dt)))
    print(i, out[f"config{i}_trajectory{i}"])

L = M.TrajectoryLoader(f"{ROOT}/data/trajectory3.npz"); T = M.BatchedTracker(L)
z3 = np.load(f"{ROOT}/data/trajectory3.npz")
N = len(z3["U"])
Ev = M.PlannerEvaluator(T, N=N, simpson_sign=+1)
z = Ev.pack(z3["X"], z3["U"], z3["S"])
for tiles in (1, 64):
    zt = torch.from_numpy(np.tile(z, (tiles, 1))).cuda()
    lam = torch.from_numpy(np.random.default_rng(7).normal(size=(tiles, N, 5))).cuda()
    o = Ev.eval_defects(zt, lam=lam, want_jac=True, want_hess=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for rep in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); Ev.eval_defects(zt, lam=lam, want_jac=True, want_hess=True, out=o); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    ms = float(np.median(ms[2:]))
    n_int = tiles * N
    byt = n_int * (5 + 60 + 144 + 5) * 8 + tiles * (8 * N + 5) * 8
    out[f"config5_planner_x{tiles}"] = dict(intervals=n_int, ms=ms, intervals_per_s=n_int / ms * 1e3, GBps=byt / ms / 1e6,
                                            hbm_frac_of_measured_6458=byt / ms / 1e6 / 6458.7)
    print(out[f"config5_planner_x{tiles}"])
json.dump(out, open(f"{ROOT}/gpurun_out/configs.json", "w"), indent=1)
