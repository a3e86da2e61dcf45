"""Multi-GPU parity on real hardware: a batch sharded over the ranks and all-gathered with NCCL is
bit-identical to the unsharded solve.  Needs at least two CUDA devices (skipped on a one-GPU box); the host logic of the
same code is covered on CPU by tests/test_sharding_gloo.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fleet(M, n=37):
    """A small fleet on trajectory1 with per-vehicle scenario constants and start states."""
    rng = np.random.default_rng(5)
    scen = [M.make_scenario(2, obs_trigger_s=float(rng.uniform(20, 60)), obs_start_s=float(rng.uniform(90, 120)),
                            obs_end_s=250.0, tl_pos=float(rng.uniform(140, 200)), tl_stop_duration=3.0) for _ in range(n)]
    xi = np.zeros((n, 5))
    xi[:, 0] = rng.uniform(0, 20, n)
    xi[:, 4] = rng.uniform(0.5, 4.0, n)
    return scen, xi


def _worker(rank, world, port, B, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        sys.path.insert(0, ROOT)
        import safe_autonomous_driving_mpc_b200 as M
        from oracle import tracker_port as P
        traj = os.path.join(ROOT, "data", "trajectory3.npz")
        T = M.BatchedTracker(M.TrajectoryLoader(traj), device=rank)
        x0, obs, n = P.monte_carlo_problems(P.RefTable.from_npz(traj), B)
        lo, hi = M.sharding.shard_bounds(B, rank, world)
        d = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev) for a in (x0, obs, n)]
        out = T.solve_batch(*d)
        U, st = M.sharding.gather_device(out["U"], out["status"], B)          # NCCL
        stats = M.sharding.reduce_stats(out["status"].cpu().numpy(), out["iters"].cpu().numpy(), 1.0 + rank, device=dev)
        # closed loop sharded by scenario: every rank drives its own vehicles; nothing to exchange but the results
        T2 = M.BatchedTracker(M.TrajectoryLoader(os.path.join(ROOT, "data", "trajectory1.npz")), device=rank)
        scen, xi = _fleet(M)
        r = M.sharding.simulate_sharded(lambda lo, hi: M.BatchedSimulation(T2, scen[lo:hi], x_init=xi[lo:hi]), len(scen), device=dev)
        q.put((rank, U.cpu().numpy(), st.cpu().numpy(), stats, r["x"], r["steps"], r["unsolved"]))
    finally:
        dist.destroy_process_group()


def test_sharded_solve_equals_unsharded_on_gpus():
    import torch
    world = min(torch.cuda.device_count(), int(os.environ.get("MPCB_TEST_WORLD", "2")))
    if world < 2:
        pytest.skip("needs at least two CUDA devices")
    import torch.multiprocessing as mp
    B = 16384 + 1                                     # ragged shards; every shard large enough for the bulk kernel
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    import safe_autonomous_driving_mpc_b200 as M
    from oracle import tracker_port as P
    traj = os.path.join(ROOT, "data", "trajectory3.npz")
    T = M.BatchedTracker(M.TrajectoryLoader(traj), device=0, coop_max_batch=0)   # shards of > 3,072 problems run the bulk kernel
    x0, obs, n = P.monte_carlo_problems(P.RefTable.from_npz(traj), B)
    ref = T.solve_batch_host(x0, obs, n)
    hist = np.bincount(ref["status"], minlength=3)
    T1 = M.BatchedTracker(M.TrajectoryLoader(os.path.join(ROOT, "data", "trajectory1.npz")), device=0)
    scen, xi = _fleet(M)
    sim = M.BatchedSimulation(T1, scen, x_init=xi)
    sim.run()
    x1, steps1, uns1 = sim.state()
    assert steps1.min() > 50
    for rank, U, st, stats, sx, ssteps, suns in got:
        assert np.array_equal(U, ref["U"]) and np.array_equal(st, ref["status"])        # bitwise: problems are independent
        assert (stats["solved"], stats["maxiter"], stats["infeasible"]) == tuple(int(v) for v in hist[:3])
        assert stats["problems"] == B and stats["ms_max"] == float(world)
        # the sharded fleet arrives exactly like the unsharded one (small shards: both run one warp per problem)
        assert np.array_equal(ssteps, steps1) and np.array_equal(suns, uns1) and np.array_equal(sx, x1)
