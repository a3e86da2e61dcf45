"""GPU parity tests of the planner evaluator (mpcb_hs_eval / mpcb_hs_nodes) through the C ABI."""
import numpy as np
import pytest

from conftest import golden, traj_path

pytestmark = pytest.mark.gpu


def _evaluator(gpu_trackers, i, N, sign=-1, **kw):
    import safe_autonomous_driving_mpc_b200 as M
    L, T = gpu_trackers[i]
    return M.PlannerEvaluator(T, N=N, simpson_sign=sign, **kw)


@pytest.mark.parametrize("i", [1, 2, 3])
def test_closure_values_match_reference(i, gpu_trackers):
    """defects (committed sign), node rows, control rows, cost against the unmodified reference's closures."""
    g = golden(f"planner_traj{i}")
    N = int(g["N"])
    E = _evaluator(gpu_trackers, i, N, s_total=float(g["s_total"]), v_max=float(g["v_max"]))
    r = E.evaluate_host(g["z"], s0=g["x0"][:, 0], want_jac=False)
    scale = np.maximum(1.0, np.abs(g["defect"]))
    assert (np.abs(r["defect"] - g["defect"]) / scale).max() <= 1e-12          # cos/sin differ by <= 1-2 ulp
    assert np.abs(r["node_rows"] - g["node_rows"]).max() <= 1e-12 * 200
    assert np.array_equal(r["ctrl_rows"], g["ctrl_rows"])
    assert (np.abs(r["cost"] - g["cost"]) / np.maximum(1.0, np.abs(g["cost"]))).max() <= 1e-13
    assert np.allclose(r["cost_terms"].sum(axis=1), r["cost"], rtol=1e-13)


@pytest.mark.parametrize("sign", [-1, 1])
def test_jacobian_and_hessian_blocks(sign, gpu_trackers, port_tables):
    """Jacobian blocks vs complex-step derivative of the oracle; Lagrangian-Hessian blocks vs central differences."""
    from oracle import planner_port as Q
    g = golden("planner_traj3")
    N = int(g["N"])
    E = _evaluator(gpu_trackers, 3, N, sign=sign)
    lam = np.random.default_rng(7).normal(size=(len(g["z"]), N, 5))             # lambda ~ N(0,1), seed 7
    r = E.evaluate_host(g["z"], lam=lam, want_jac=True, want_hess=True)
    tab = port_tables[3]
    for w in range(1, len(g["z"]), 2):                                         # perturbed windows (off-knot)
        X, U, _ = Q.unpack(g["z"][w], N)
        for k in (0, 4, 7, 11):
            Jr = Q.hs_defect_jac(tab, X[k], X[k + 1], U[k], simpson_sign=sign)
            assert np.abs(r["jac"][w, k] - Jr).max() <= 1e-11
            Hr = Q.hs_lagrangian_hess(tab, X[k], X[k + 1], U[k], lam[w, k], simpson_sign=sign)
            H = r["hess"][w, k]
            assert np.abs(H - Hr).max() <= 1e-7 * max(1.0, np.abs(Hr).max())
            assert np.array_equal(H, H.T)
    d_ref = np.array([Q.defects(tab, z, N, simpson_sign=sign) for z in g["z"]])
    assert (np.abs(r["defect"] - d_ref) / np.maximum(1.0, np.abs(d_ref))).max() <= 1e-12


def test_cost_gradient(gpu_trackers):
    from oracle import planner_port as Q
    g = golden("planner_traj2")
    N = int(g["N"])
    s_total = float(g["s_total"])
    E = _evaluator(gpu_trackers, 2, N, s_total=s_total)
    r = E.evaluate_host(g["z"][:4], s0=g["x0"][:4, 0], want_jac=False)
    for w in range(4):
        z = g["z"][w]
        gr = np.zeros_like(z)
        for j in range(len(z)):
            e = np.zeros_like(z)
            e[j] = 1e-6
            gr[j] = (Q.cost(z + e, N, g["x0"][w], s_total) - Q.cost(z - e, N, g["x0"][w], s_total)) / 2e-6
        assert np.abs(r["cost_grad"][w] - gr).max() <= 1e-6 * max(1.0, np.abs(gr).max())


def test_whole_trajectory_as_one_batch_and_tiled(gpu_trackers, port_tables):
    """BASELINE config 5: trajectory3's 1,258 intervals as one batch (sign +1: s, k, v defects vanish on the committed
    data), and the x64 tiling (80,512 intervals): every tile bitwise equal to the first (determinism, no cross-talk)."""
    import torch
    z3 = np.load(traj_path(3))
    X, U, S = z3["X"], z3["U"], z3["S"]
    N = len(U)
    E = _evaluator(gpu_trackers, 3, N, sign=+1)
    z = E.pack(X, U, S)
    r1 = E.evaluate_host(z, want_jac=True)
    assert np.abs(r1["defect"][0][:, [3, 4]]).max() < 1e-6
    assert np.median(np.abs(r1["defect"][0][:, 0])) < 1e-5
    zt = torch.from_numpy(np.tile(z, (64, 1))).cuda()
    lam = torch.from_numpy(np.random.default_rng(7).normal(size=(64, N, 5))).cuda()
    lam[:] = lam[0]
    out = E.eval_defects(zt, lam=lam, want_jac=True, want_hess=True)
    torch.cuda.synchronize()
    for k in ("defect", "jac", "hess"):
        a = out[k].cpu().numpy()
        assert np.array_equal(a[0], a[63]) and np.array_equal(a[0], a[31])
    # the values-only / with-derivatives kernels are separate instantiations (different FMA contraction): ulp-level
    assert np.abs(out["defect"][0].cpu().numpy() - r1["defect"][0]).max() <= 1e-13
    assert np.abs(out["jac"][0].cpu().numpy() - r1["jac"][0]).max() <= 1e-13
    nodes = E.eval_nodes(zt, zt[:, 0].contiguous())
    torch.cuda.synchronize()
    assert torch.isfinite(nodes["cost"]).all() and torch.equal(nodes["cost"][0], nodes["cost"][63])
