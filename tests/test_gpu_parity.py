"""GPU parity tests proper: everything goes through the C ABI (ctypes -> libmpcb200.so -> sm_100a kernels) and is
compared with vectors produced by the unmodified reference (tests/golden, see tools/make_golden.py) or with the
CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): optimal control sequence within 1e-4 absolute, objective within 1e-6 relative
(absolute floor 1e-6: J is ~1e-4 in steady tracking), identical feasibility flags and active sets.
Function values: 1e-12 (fp64 with FMA contraction on the device vs numpy without)."""
import numpy as np
import pytest

from conftest import golden, traj_path

pytestmark = pytest.mark.gpu

U_TOL = 1e-4
J_RTOL = 1e-6
FN_TOL = 1e-12

SETS = [("solve_traj1", 1), ("solve_traj2", 2), ("solve_traj3", 3), ("solve_mc_traj3", 3)]


@pytest.mark.parametrize("i", [1, 2, 3])
def test_predict_cost_constraints(i, gpu_trackers):
    """K1 values vs reference predict / cost / constraints_wrapper (trajectory_tracking.py:87-211)."""
    z = golden(f"fn_traj{i}")
    _, T = gpu_trackers[i]
    r = T.eval_batch(z["x0"], z["U"], z["obs_sv"], z["n_obs"])
    scale = np.maximum(1.0, np.abs(z["predict"]))
    assert np.max(np.abs(r["Xpred"] - z["predict"]) / scale) < FN_TOL
    assert np.max(np.abs(r["cost"] - z["cost"]) / np.maximum(1.0, np.abs(z["cost"]))) < FN_TOL
    assert np.array_equal(np.isnan(r["cons"]), np.isnan(z["constraints"]))
    m = ~np.isnan(z["constraints"])
    assert np.max(np.abs(r["cons"][m] - z["constraints"][m]) / np.maximum(1.0, np.abs(z["constraints"][m]))) < FN_TOL


@pytest.mark.parametrize("name,i", SETS)
def test_warm_start(name, i, gpu_trackers):
    """trajectory_tracking.py:223-246, including the latched braking guess and the s >= s_max branch."""
    z = golden(name)
    _, T = gpu_trackers[i]
    w = T.eval_batch(z["x0"], np.zeros((len(z["x0"]), 10)), z["obs_sv"], z["n_obs"])["warm"]
    assert np.max(np.abs(w - z["U_init"])) < FN_TOL


@pytest.mark.parametrize("i", [1, 2, 3])
def test_linearisation_against_oracle_model(i, gpu_trackers, port_tables):
    """Exact derivatives on the device vs the numpy model (itself checked against central differences of the
    reference functions in test_algorithm_model.py)."""
    from oracle import sqp_admm_model as A
    z = golden(f"fn_traj{i}")
    _, T = gpu_trackers[i]
    r = T.eval_batch(z["x0"], z["U"], z["obs_sv"], z["n_obs"])["lin"]
    asm = A.assemble(port_tables[i], z["x0"], z["U"])
    qp = A.build_qp(asm, z["x0"], z["U"], z["obs_sv"], z["n_obs"])
    H = np.array([[qp["P"][b][a, c] for a in range(10) for c in range(a + 1)] for b in range(len(z["x0"]))])
    assert np.max(np.abs(r[:, :55] - H)) < 1e-9 * max(1.0, np.abs(H).max())
    assert np.max(np.abs(r[:, 55:65] - qp["q"])) < 1e-9 * max(1.0, np.abs(qp["q"]).max())
    assert np.max(np.abs(r[:, 65:105].reshape(-1, 4, 10) - asm["dX"][:, 2:6, 1])) < 1e-12
    assert np.max(np.abs(r[:, 105:145].reshape(-1, 4, 10) - asm["dX"][:, 2:6, 2])) < 1e-12


def _active_rows(c, n_obs, tol):
    return c[: 5 * (7 + n_obs)] <= tol


def _other_local_optimum(tab, x0, obs, U_ours, J_ours, J_ref):
    """The tracking NLP is nonconvex (bilinear dynamics, table looked up at the predicted s): in the tight bend of
    trajectory3 (s ~ 720 m) a few problems have TWO local optima, and which one a solver reaches depends on its path
    (SLSQP damps its steps with a line search, the Gauss-Newton SQP here takes full steps).  A returned point counts as
    'another local optimum' only if the ORACLE's own SLSQP, started from it at tight tolerance, stays there (so it is
    a stationary point of the reference formulation, not a solver artefact), it is feasible under the oracle's
    constraints, and its objective is within 5 % of the reference's."""
    from oracle import tracker_port as P
    if P.constraint_values(tab, U_ours, x0, obs).min() < -1e-6:
        return False
    r = P.solve_converged(tab, x0, obs, U_start=U_ours)
    return (np.max(np.abs(r.x - U_ours)) <= U_TOL and r.status in (0, 8) and abs(J_ours - J_ref) <= 0.05 * max(1.0, abs(J_ref)))


@pytest.mark.parametrize("variant", ["default", "bulk", "bulk_thread2"])
@pytest.mark.parametrize("name,i", SETS)
def test_solve_against_converged_reference(name, i, variant, tracker_variants, port_tables):
    """Solve-level parity against the reference formulation solved to convergence, through every execution
    shape: the warp-per-problem kernel a small batch gets by default, and -- forced with coop_max_batch=0 -- the
    thread-per-problem first pass that carries the 65,536-problem benchmark (with either robust pass behind it)."""
    from oracle import tracker_port as P
    z = golden(name)
    _, T = tracker_variants[variant][i]
    s = T.solve_batch_host(z["x0"], z["obs_sv"], z["n_obs"])
    U = s["U"].reshape(-1, 10)
    pin = z["pinned"].copy()
    assert pin.sum() >= 0.8 * len(pin)
    # --- pinned problems: U, objective, flags, active set -------------------------------------
    err = np.abs(U - z["U_conv"]).max(axis=1)
    # problems with two local optima (see _other_local_optimum): at most 0.1 % of a set, each verified with the oracle;
    # they keep every other check below except the comparison with the reference's optimum
    other = []
    for b in np.where(pin & (err > U_TOL))[0]:
        obs = [tuple(o) for o in z["obs_sv"][b, : z["n_obs"][b]]]
        assert s["status"][b] == 0 and _other_local_optimum(port_tables[i], z["x0"][b], obs, U[b], s["obj"][b], z["J_conv"][b]), \
            f"problem {b}: |dU| = {err[b]:.3e} and the returned point is not a stationary point of the reference NLP"
        other.append(int(b))
    assert len(other) <= max(1, len(pin) // 1000), other
    pin[other] = False
    assert err[pin].max() <= U_TOL, f"max |dU| = {err[pin].max():.3e} at {np.argmax(np.where(pin, err, 0))}"
    jerr = np.abs(s["obj"] - z["J_conv"]) / np.maximum(np.abs(z["J_conv"]), 1.0)
    assert jerr[pin].max() <= J_RTOL
    assert np.all(s["status"][pin] == 0), "pinned (feasible, converged) problems must come back SOLVED"
    assert np.all(s["cmin"][pin] >= -1e-6)
    if variant != "default":
        assert T.last_pass_ms()[0] > 0.0 and not T.last_call_used_coop(), "the thread-per-problem first pass must have run"
    n_amb = 0
    for b in np.where(pin)[0]:
        no = int(z["n_obs"][b])
        c_ref = z["c_conv"][b]
        ours = np.array([(int(s["active"][b]) >> r) & 1 for r in range(5 * (7 + no))], dtype=bool)
        ref_active = _active_rows(c_ref, no, 1e-6)
        ref_clear = c_ref[: 5 * (7 + no)] > 1e-3
        assert np.all(ours[ref_active]), f"problem {b}: reference-active row inactive on the GPU"
        assert not np.any(ours[ref_clear]), f"problem {b}: GPU-active row is clearly inactive in the reference"
        n_amb += int(np.sum(~ref_active & ~ref_clear))
        # bound activity
        lb, ub = np.tile(P.U_MIN, 5), np.tile(P.U_MAX, 5)
        ref_b = (z["U_conv"][b] - lb <= 1e-6) | (ub - z["U_conv"][b] <= 1e-6)
        ref_b_clear = (z["U_conv"][b] - lb > 1e-3) & (ub - z["U_conv"][b] > 1e-3)
        ours_b = np.array([(int(s["active"][b]) >> (45 + k)) & 1 for k in range(10)], dtype=bool)
        assert np.all(ours_b[ref_b]) and not np.any(ours_b[ref_b_clear])
    assert n_amb <= 0.02 * pin.sum() * 45
    # --- predicted trajectory is the reference's predict() at our U ------------------------------
    for b in np.where(pin)[0][:16]:
        Xp = P.predict(port_tables[i], z["x0"][b], U[b])
        assert np.max(np.abs(Xp - s["Xpred"][b]) / np.maximum(1.0, np.abs(Xp))) < FN_TOL
        assert s["obj"][b] == pytest.approx(P.cost(port_tables[i], U[b], z["x0"][b]), rel=1e-12, abs=1e-12)
    # --- unpinned problems: feasibility flags only ---------------------------------------------
    # Where the reference's SLSQP ended on an infeasible point we must flag INFEASIBLE -- unless the point we
    # return is feasible under the reference's own constraint function, which proves the problem feasible and the
    # reference's exit a solver failure (mode 8 exits); those are counted, not hidden.
    infeas_ref = np.where((~pin) & (z["min_c"] < -1e-5))[0]
    proven_feasible = 0
    for b in infeas_ref:
        if s["status"][b] == 2:
            continue
        obs = [tuple(o) for o in z["obs_sv"][b, : z["n_obs"][b]]]
        c = P.constraint_values(port_tables[i], U[b], z["x0"][b], obs)
        assert c.min() >= -1e-6, f"problem {b}: reference infeasible, GPU says feasible but its point is not"
        proven_feasible += 1
    assert proven_feasible <= 0.05 * len(pin)


def test_reference_api_solve_signature(gpu_trackers):
    """BatchedTracker.solve has the reference's call surface (trajectory_tracking.py:213-263)."""
    z = golden("solve_traj2")
    _, T = gpu_trackers[2]
    b = int(np.where(z["pinned"] & (z["n_obs"] == 1))[0][0])
    obs = [{"s": float(z["obs_sv"][b, 0, 0]), "v": float(z["obs_sv"][b, 0, 1]), "type": "car"}]
    u0, pred, sec = T.solve(z["x0"][b], obs)
    assert u0.shape == (2,) and pred.shape == (6, 5) and isinstance(sec, float)
    assert np.max(np.abs(u0 - z["U_conv"][b, :2])) <= U_TOL
    assert np.array_equal(pred[0], z["x0"][b])
    assert T.dt == 0.2 and T.N == 5 and T.obstacle_safety_distance == 5.0
    assert np.array_equal(T.u_min, [-0.6, -5.0]) and np.array_equal(T.u_max, [0.6, 4.0])
    assert np.array_equal(T.dynamics(np.array([1.0, 2, 3, 4, 5]), np.array([6.0, 7]), 0.5), [5, 15, 17.5, 6, 7])


def test_edge_cases(gpu_trackers):
    """Empty batch, ragged obstacle counts, states past the end of the table, zero speed, out-of-range n_obs."""
    L, T = gpu_trackers[1]
    r = T.solve_batch_host(np.zeros((0, 5)), np.zeros((0, 2, 2)), np.zeros(0, np.int32))
    assert r["U"].shape == (0, 5, 2)
    x0 = np.array([[L.s_max + 3.0, 0.0, 0.0, 0.0, 2.0],      # past the end: get_state returns the last row
                   [L.s_max - 0.5, 0.01, 0.0, 0.0, 3.0],      # crosses s_max inside the horizon
                   [10.0, 0.0, 0.0, 0.0, 0.0],                # standing still
                   [-5.0, 0.0, 0.0, 0.0, 1.0],                # before the first knot (extrapolation)
                   [100.0, 0.0, 0.0, 0.0, 5.0]])
    obs = np.zeros((5, 2, 2)); obs[4, 0] = (130.0, 2.0)
    n = np.array([0, 0, 0, 0, 7], dtype=np.int32)             # 7 is clamped to 2 (second obstacle: s=0 -> behind)
    r = T.solve_batch_host(x0, obs, n)
    assert np.all(np.isfinite(r["U"])) and np.all(np.isfinite(r["obj"]))
    assert np.all(r["U"][:, :, 0] >= -0.6 - 1e-9) and np.all(r["U"][:, :, 0] <= 0.6 + 1e-9)
    assert np.all(r["U"][:, :, 1] >= -5.0 - 1e-9) and np.all(r["U"][:, :, 1] <= 4.0 + 1e-9)
    from oracle import tracker_port as P
    tab = P.RefTable.from_npz(traj_path(1))
    for b in range(4):
        assert r["obj"][b] == pytest.approx(P.cost(tab, r["U"][b].ravel(), x0[b]), rel=1e-12, abs=1e-12)


def test_full_size_batch_properties(gpu_trackers, port_tables):
    """BASELINE config 4 at full size (65,536 problems on trajectory3): size-independent properties, judged with the
    oracle's functions only (tests/kkt.py) -- nothing the device reports about itself is trusted except the controls.
    (a) shard/permutation invariance is bitwise (problems are independent);
    (b) every SOLVED problem satisfies every row of the reference's constraints_wrapper, evaluated by the oracle at the
        returned controls, and the bounds; every INFEASIBLE flag comes with a violated oracle row;
    (c) u1_4 == 0 at every optimum (SURVEY A.1 known answer);
    (d) first-order optimality WITH multipliers: for every solved problem a non-negative least-squares fit of the
        oracle's cost gradient on the gradients of its active rows and bounds leaves no residual;
    (e) objective, cmin and the active-set bits returned equal the oracle's values at the returned controls."""
    from oracle import tracker_port as P
    import kkt
    tab = port_tables[3]
    _, T = gpu_trackers[3]
    B = 65536
    x0, obs, n = P.monte_carlo_problems(tab, B)
    full = {k: v.copy() for k, v in T.solve_batch_host(x0, obs, n).items()}
    assert not T.last_call_used_coop()
    # (a) two shards, then a permutation
    h = B // 2
    for sl in (slice(0, h), slice(h, B)):
        part = T.solve_batch_host(x0[sl], obs[sl], n[sl])
        for k in ("U", "Xpred", "obj", "status", "iters", "cmin", "active"):
            assert np.array_equal(part[k], full[k][sl]), f"shard result differs in {k}"
    perm = np.random.default_rng(5).permutation(B)[:8192]
    part = T.solve_batch_host(x0[perm], obs[perm], n[perm])
    assert np.array_equal(part["U"], full["U"][perm]) and np.array_equal(part["status"], full["status"][perm])
    # (b) feasibility from the oracle's rows
    U = full["U"].reshape(B, 10)
    c, G, asm, G2, kink = kkt.reference_rows(tab, x0, U, obs, n, with_kinks=True)
    cmin = np.nanmin(c, axis=1)
    ok = full["status"] == 0
    assert ok.mean() > 0.9
    assert cmin[ok].min() >= -1e-6, f"a SOLVED problem violates a reference row by {cmin[ok].min():.3e}"
    assert np.all(U >= np.tile(P.U_MIN, 5) - 1e-12) and np.all(U <= np.tile(P.U_MAX, 5) + 1e-12)
    bad = full["status"] == 2
    assert np.all(cmin[bad] < -1e-6), "an INFEASIBLE flag on a point that satisfies every reference row"
    assert np.all(full["status"] != 1), "no problem of the Monte-Carlo set should run out of iterations"
    # (c)
    assert np.max(np.abs(U[ok, 8])) < 1e-6
    # (e) what the device reports equals the oracle's evaluation at the same controls
    assert np.max(np.abs(asm["cost"] - full["obj"]) / np.maximum(1.0, np.abs(asm["cost"]))) < 1e-12
    assert np.max(np.abs(cmin - full["cmin"]) / np.maximum(1.0, np.abs(cmin))) < 1e-10
    bits = ((full["active"][:, None] >> np.arange(45, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    valid = ~np.isnan(c)
    clear_on = valid & (c <= 1e-6 - 1e-9)
    clear_off = valid & (c > 1e-6 + 1e-9)
    assert np.all(bits[clear_on]) and not np.any(bits[clear_off]) and not np.any(bits[~valid])
    # (d) KKT with multipliers, all solved problems
    g = kkt.cost_gradient(asm, U)
    res, nact = kkt.kkt_residual(c[ok], G[ok], g[ok], U[ok], G2=G2[ok], kink=kink[ok])
    assert (nact > 0).sum() > 500, "the set must contain problems that sit on constraints"
    assert res.max() <= 1e-5, f"KKT residual {res.max():.3e} at solved problem {np.where(ok)[0][np.argmax(res)]}"


def test_device_tensor_entry_point(gpu_trackers, port_tables):
    """mpcb_solve_batch with device pointers (torch used for memory + stream hand-off only) equals the host path."""
    import torch
    from oracle import tracker_port as P
    _, T = gpu_trackers[3]
    x0, obs, n = P.monte_carlo_problems(port_tables[3], 65536)
    x0, obs, n = x0[:4096], obs[:4096], n[:4096]
    host = {k: v.copy() for k, v in T.solve_batch_host(x0, obs, n).items()}
    dev = torch.device("cuda:0")
    out = T.solve_batch(torch.from_numpy(x0).to(dev), torch.from_numpy(obs).to(dev), torch.from_numpy(n).to(dev))
    torch.cuda.synchronize()
    assert np.array_equal(out["U"].cpu().numpy(), host["U"])
    assert np.array_equal(out["status"].cpu().numpy(), host["status"])
    assert np.array_equal(out["obj"].cpu().numpy(), host["obj"])
    assert T.launch_count() > 0 and T.last_kernel_ms() > 0.0


def test_execution_shapes_agree(gpu_trackers, port_tables):
    """The same problems through every execution shape -- two-pass with the warp-per-problem robust pass (default),
    two-pass with the thread-per-problem robust pass, single robust pass, and the small-batch warp-per-problem path --
    give the same flags and the same controls to 1e-6 (the shapes differ in summation order only)."""
    import safe_autonomous_driving_mpc_b200 as M
    from oracle import tracker_port as P
    L, T = gpu_trackers[3]
    x0, obs, n = P.monte_carlo_problems(port_tables[3], 6000)
    ref = {k: v.copy() for k, v in T.solve_batch_host(x0, obs, n).items()}
    variants = dict(thread_pass2=dict(coop_pass2=0), robust_only=dict(fast_pass=0))
    for name, kw in variants.items():
        Tv = M.BatchedTracker(L, **kw)
        r = Tv.solve_batch_host(x0, obs, n)
        agree = r["status"] == ref["status"]
        assert agree.mean() > 0.999, name                     # borderline flags may flip between solver paths
        ok = agree & (ref["status"] == 0)
        assert np.abs(r["U"] - ref["U"])[ok].max() <= 1e-6, name
    assert not T.last_call_used_coop()
    for nb in (700, 3000):                                      # B <= coop_max_batch: warp-per-problem first pass
        small = T.solve_batch_host(x0[:nb], obs[:nb], n[:nb])   # (700: packed staging path, 3000: chunked copies)
        assert T.last_call_used_coop()
        same = small["status"] == ref["status"][:nb]
        assert same.mean() > 0.999
        ok = same & (small["status"] == 0)
        assert np.abs(small["U"] - ref["U"][:nb])[ok].max() <= 1e-6


def test_argument_errors_and_degenerate_sizes(gpu_trackers):
    import ctypes as C
    from safe_autonomous_driving_mpc_b200 import _lib
    L, T = gpu_trackers[1]
    lib = T._lib
    assert lib.mpcb_solve_batch_host(T._h, 0, None, None, None, None, None, None, None, None, None, None) == 0   # B = 0
    assert lib.mpcb_solve_batch_host(T._h, -1, None, None, None, None, None, None, None, None, None, None) == -1
    assert lib.mpcb_solve_batch_host(T._h, 4, None, None, None, None, None, None, None, None, None, None) == -1  # null inputs
    assert lib.mpcb_solve_batch_host(None, 1, None, None, None, None, None, None, None, None, None, None) == -1
    x0 = np.array([[10.0, 0.0, 0.0, 0.0, 5.0]] * 3)
    obs = np.zeros((3, 2, 2))
    # n_obs outside 0..2 is clamped like the reference's list slicing would
    r = T.solve_batch_host(x0, obs, np.array([0, -3, 0], dtype=np.int32))
    assert np.array_equal(r["U"][0], r["U"][1])
    # states at and beyond the end of the table (get_state clamps to the last row, trajectory_loader.py:90-91)
    xe = np.array([[L.s_max - 0.5, 0.0, 0.0, 0.0, 1.0], [L.s_max + 3.0, 0.0, 0.0, 0.0, 1.0]])
    r = T.solve_batch_host(xe, np.zeros((2, 2, 2)), np.zeros(2, dtype=np.int32))
    assert np.all(np.isfinite(r["U"])) and np.all(np.isfinite(r["obj"]))
    # device entry point: rows are written with 128-bit stores, so U_out / Xpred_out must be 16-byte aligned
    import torch
    dx = torch.from_numpy(x0).cuda(); do = torch.from_numpy(obs).cuda(); dn = torch.zeros(3, dtype=torch.int32, device="cuda")
    buf = torch.zeros(3 * 10 + 2, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ok = lib.mpcb_solve_batch(T._h, 3, dx.data_ptr(), do.data_ptr(), dn.data_ptr(), buf.data_ptr(), None, None, None, None, None, None, C.c_void_p(st))
    assert ok == 0
    bad = lib.mpcb_solve_batch(T._h, 3, dx.data_ptr(), do.data_ptr(), dn.data_ptr(), buf.data_ptr() + 8, None, None, None, None, None, None, C.c_void_p(st))
    assert bad == -1
    torch.cuda.synchronize()
    assert np.array_equal(buf[:30].cpu().numpy().reshape(3, 5, 2), T.solve_batch_host(x0, obs, np.zeros(3, dtype=np.int32))["U"])
    p = _lib.Params()
    lib.mpcb_default_params(C.byref(p))
    p.N = 7
    h = C.c_void_p()
    assert lib.mpcb_create(C.byref(h), C.byref(p), L._h, 0) == -4          # MPCB_ERR_UNSUPPORTED
    p.N = 5
    p.alpha = 2.5
    assert lib.mpcb_create(C.byref(h), C.byref(p), L._h, 0) == -1
    assert lib.mpcb_create(C.byref(h), C.byref(p), L._h, 99) != 0


def test_repeated_host_calls_replay_fresh_data(gpu_trackers, port_tables):
    """mpcb_solve_batch_host captures a CUDA graph on the second call with the same buffers and replays it afterwards:
    the replays must read the buffers' CURRENT contents (both the chunked large-batch path, which copies straight
    from / to the caller's page-locked arrays, and the packed small-batch path), pageable buffers must keep working
    (no graph), and the answers must equal those of a fresh handle."""
    import safe_autonomous_driving_mpc_b200 as M
    from oracle import tracker_port as P
    L, _ = gpu_trackers[3]
    x0, obs, n = P.monte_carlo_problems(port_tables[3], 3 * 6000)
    for B in (6000, 700):                                    # chunked path / packed path
        T = M.BatchedTracker(L)
        fresh = M.BatchedTracker(L)
        pin = {k: M.tracker.PinnedBuffer(a[:B].shape, a.dtype) for k, a in (("x0", x0), ("obs", obs), ("n", n))}
        for rep in range(5):                                 # 0: direct, 1: capture + replay, 2..: replay
            sl = slice((rep % 3) * B, (rep % 3) * B + B)
            pin["x0"].array[...] = x0[sl]; pin["obs"].array[...] = obs[sl]; pin["n"].array[...] = n[sl]
            r = {k: v.copy() for k, v in T.solve_batch_host(pin["x0"].array, pin["obs"].array, pin["n"].array).items()}
            ref = fresh.solve_batch_host(x0[sl].copy(), obs[sl].copy(), n[sl].copy(), pinned_out=False)   # pageable: never a graph
            assert np.array_equal(r["status"], ref["status"]), (B, rep)
            assert np.array_equal(r["U"], ref["U"]), (B, rep)
            assert np.array_equal(r["Xpred"], ref["Xpred"]), (B, rep)
            assert np.array_equal(r["active"], ref["active"]), (B, rep)


def test_first_pass_caps_do_not_change_answers(gpu_trackers, port_tables):
    """The caps of the two-level policy in the thread-per-problem kernel only decide WHERE a problem is solved (first
    pass or robust pass), not what comes out: flags equal, controls equal to 1e-6 across cap settings."""
    import safe_autonomous_driving_mpc_b200 as M
    from oracle import tracker_port as P
    L, T = gpu_trackers[3]
    x0, obs, n = P.monte_carlo_problems(port_tables[3], 6000)
    ref = {k: v.copy() for k, v in T.solve_batch_host(x0, obs, n).items()}
    for kw in (dict(thread_max_rounds=4, thread_max_segments=2, thread_fail_rounds=0),
               dict(thread_max_rounds=6, thread_max_segments=4, thread_fail_rounds=0),
               dict(thread_max_rounds=5, thread_max_segments=1, thread_fail_rounds=5)):
        Tv = M.BatchedTracker(L, **kw)                         # kept alive: the results are views of ITS pinned buffers
        r = Tv.solve_batch_host(x0, obs, n)
        agree = r["status"] == ref["status"]
        assert agree.mean() > 0.999, kw
        ok = agree & (ref["status"] == 0)
        assert np.abs(r["U"] - ref["U"])[ok].max() <= 1e-6, kw


def test_timing_queries_after_graph_replay(gpu_trackers, port_tables):
    """mpcb_last_kernel_ms / mpcb_last_pass_ms after a REPLAYED host call (the third identical call onwards runs as a
    CUDA graph; its timing events are external event-record nodes of that graph) -- packed and chunked paths."""
    import safe_autonomous_driving_mpc_b200 as M
    from oracle import tracker_port as P
    L, _ = gpu_trackers[3]
    x0, obs, n = P.monte_carlo_problems(port_tables[3], 6000)
    for B in (1, 700, 6000):
        T = M.BatchedTracker(L)
        pin = {k: M.tracker.PinnedBuffer(a[:B].shape, a.dtype) for k, a in (("x0", x0), ("obs", obs), ("n", n))}
        pin["x0"].array[...] = x0[:B]; pin["obs"].array[...] = obs[:B]; pin["n"].array[...] = n[:B]
        for rep in range(5):
            T.solve_batch_host(pin["x0"].array, pin["obs"].array, pin["n"].array)
            ms = T.last_kernel_ms()
            assert 0.0 < ms < 50.0, (B, rep, ms)
            if B <= 2048:                                     # packed path keeps the per-pass split, replayed or not
                a, b, k = T.last_pass_ms()
                assert a > 0.0 and b >= 0.0 and 0 <= k <= B, (B, rep, a, b, k)
    # the reference-shaped call: solve() three times, then the timing query
    T = M.BatchedTracker(L)
    for _ in range(4):
        T.solve(x0[0], [])
    assert T.last_kernel_ms() > 0.0 and T.last_pass_ms()[0] > 0.0


def test_closed_loop_form_of_the_host_entry(gpu_trackers, port_tables):
    """mpcb_solve_batch_host_u0 returns exactly U*[0], status and objective of mpcb_solve_batch_host (same solve, 20-28
    bytes per problem over PCIe), for the packed and the chunked path, replayed calls included."""
    from oracle import tracker_port as P
    _, T = gpu_trackers[3]
    x0, obs, n = P.monte_carlo_problems(port_tables[3], 9000)
    for B in (1, 900, 9000):
        full = {k: v.copy() for k, v in T.solve_batch_host(x0[:B], obs[:B], n[:B]).items()}
        for rep in range(4):
            r = T.solve_batch_host_u0(x0[:B], obs[:B], n[:B], want_obj=True)
            assert np.array_equal(r["u0"], full["U"][:, 0, :]), (B, rep)
            assert np.array_equal(r["status"], full["status"]) and np.array_equal(r["obj"], full["obj"]), (B, rep)
        r = T.solve_batch_host_u0(x0[:B], obs[:B], n[:B], pinned_out=False)
        assert np.array_equal(r["u0"], full["U"][:, 0, :]) and "obj" not in r


def test_asynchronous_host_entry_with_batches_in_flight(gpu_trackers, port_tables):
    """mpcb_solve_batch_host_async / mpcb_wait with three handles in flight (the form bench.py's e2e arm uses): every
    batch comes back bit-identical to the synchronous call on the same problems -- every output and the closed-loop form
    (U*[0] + status), first calls, captured call and graph replays, with the input buffers rewritten between rounds."""
    import safe_autonomous_driving_mpc_b200 as M
    from oracle import tracker_port as P
    L, T = gpu_trackers[3]
    B, nh = 9000, 3
    x0, obs, n = P.monte_carlo_problems(port_tables[3], 2 * nh * B)
    PB = M.tracker.PinnedBuffer
    Ts = [M.BatchedTracker(L) for _ in range(nh)]
    full_spec = dict(U=((B, 5, 2), np.float64), Xpred=((B, 6, 5), np.float64), obj=((B,), np.float64),
                     status=((B,), np.int32), iters=((B, 2), np.int32), cmin=((B,), np.float64), active=((B,), np.uint64))
    cl_spec = dict(u0=((B, 2), np.float64), status=((B,), np.int32))
    refs = [{k: v.copy() for k, v in T.solve_batch_host(x0[j * B:(j + 1) * B], obs[j * B:(j + 1) * B], n[j * B:(j + 1) * B]).items()}
            for j in range(2 * nh)]
    for spec in (full_spec, cl_spec):
        bufs = [dict(x0=PB((B, 5), np.float64), obs=PB((B, 2, 2), np.float64), n=PB((B,), np.int32)) for _ in range(nh)]
        outs = [{k: PB(shp, dt) for k, (shp, dt) in spec.items()} for _ in range(nh)]
        for rnd in range(4):                                   # 0 direct, 1 captured, 2.. replayed; data alternates
            js = [(rnd % 2) * nh + k for k in range(nh)]
            for k, j in enumerate(js):
                bufs[k]["x0"].array[...] = x0[j * B:(j + 1) * B]
                bufs[k]["obs"].array[...] = obs[j * B:(j + 1) * B]
                bufs[k]["n"].array[...] = n[j * B:(j + 1) * B]
                for o in outs[k].values():
                    o.array[...] = 0
                Ts[k].solve_batch_host_async(bufs[k]["x0"].array, bufs[k]["obs"].array, bufs[k]["n"].array,
                                             {kk: o.array for kk, o in outs[k].items()})
            for k, j in enumerate(js):
                Ts[k].wait()
                for kk, o in outs[k].items():
                    want = refs[j]["U"][:, 0, :] if kk == "u0" else refs[j][kk]
                    assert np.array_equal(o.array, want), (kk, rnd, k)
    # small batches complete inside the call (packed staging path)
    o = dict(u0=np.zeros((5, 2)), status=np.zeros(5, np.int32))
    Ts[0].solve_batch_host_async(x0[:5].copy(), obs[:5].copy(), n[:5].copy(), o)
    small = T.solve_batch_host(x0[:5], obs[:5], n[:5])        # same execution shape (one warp per problem)
    assert np.array_equal(o["u0"], small["U"][:, 0, :]) and np.array_equal(o["status"], small["status"])
    Ts[0].wait()
