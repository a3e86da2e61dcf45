"""CPU tests of the planner oracle (oracle/planner_port.py) against the reference's golden vectors, and of the
device formulas compiled for the host (tests/hostbuild) against the oracle.  No GPU."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import golden, traj_path

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostbuild"))


@pytest.fixture(scope="module")
def hostlib():
    import build as hb
    return C.CDLL(hb.build())


@pytest.mark.parametrize("i", [1, 2, 3])
def test_port_matches_reference_closures(i, port_tables):
    """every closure value and the cost, as returned by the unmodified reference (tools/make_golden_planner.py)"""
    from oracle import planner_port as Q
    g = golden(f"planner_traj{i}")
    tab, N = port_tables[i], int(g["N"])
    for w in range(len(g["z"])):
        z = g["z"][w]
        assert np.array_equal(Q.defects(tab, z, N), g["defect"][w])
        assert np.array_equal(Q.node_rows(z, N, 0.0, float(g["v_max"])), g["node_rows"][w])
        assert np.array_equal(Q.ctrl_rows(z, N), g["ctrl_rows"][w])
        assert Q.cost(z, N, g["x0"][w], float(g["s_total"])) == g["cost"][w]
        X, _, _ = Q.unpack(z, N)
        assert np.array_equal(X[0] - g["x0"][w], g["initial"][w])


@pytest.mark.parametrize("i", [1, 2, 3])
def test_committed_trajectories_satisfy_plus_sign(i, port_tables):
    """the committed JSON trajectories obey x_pred = x_k + dt/6(...) in the s, k, v components (which do
    not depend on the missing k_ref spline beyond 1 - d k_ref); the committed code's sign gives metres of defect."""
    from oracle import planner_port as Q
    tab = port_tables[i]
    z = np.load(traj_path(i))
    X, U = z["X"], z["U"]
    n = len(U)
    plus = np.array([Q.hs_defect(tab, X[k], X[k + 1], U[k], simpson_sign=+1) for k in range(n)])
    minus = np.array([Q.hs_defect(tab, X[k], X[k + 1], U[k], simpson_sign=-1) for k in range(n)])
    assert np.abs(plus[:, [3, 4]]).max() < 1e-6
    assert np.median(np.abs(plus[:, 0])) < 1e-5     # s couples to the (missing) k_ref spline through 1 - d k_ref:
    assert np.abs(plus[:, 0]).max() < 1.0           # tiny on straights, up to ~0.5 m in bends with the proxy k_ref
    assert np.abs(minus[:, 0]).max() > 5.0 and np.abs(minus[:, 4]).max() > 2.0


@pytest.mark.parametrize("sign", [-1, 1])
def test_device_formulas_on_host_match_oracle(sign, hostlib, port_tables):
    """hs_interval (the exact code the kernel runs) compiled with g++: values bit-exact with the oracle, Jacobian
    against the complex-step derivative, Lagrangian Hessian against central differences of that Jacobian."""
    from oracle import planner_port as Q
    vp = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    for i in (1, 3):
        g = golden(f"planner_traj{i}")
        tab, N = port_tables[i], int(g["N"])
        s = np.ascontiguousarray(tab.s)
        y = np.ascontiguousarray(tab.X[:, 1:5])
        for w in range(0, len(g["z"]), 3):
            X, U, _ = Q.unpack(g["z"][w], N)
            xk, xn, u = np.ascontiguousarray(X[:-1]), np.ascontiguousarray(X[1:]), np.ascontiguousarray(U)
            lam = np.random.default_rng(w).normal(size=(N, 5))
            d, J, H = np.zeros((N, 5)), np.zeros((N, 5, 12)), np.zeros((N, 12, 12))
            hostlib.host_hs_eval(vp(s), vp(y), C.c_int(tab.K), C.c_double(0.3), C.c_int(sign), C.c_int(N), vp(xk),
                                 vp(xn), vp(u), vp(lam), vp(d), vp(J), vp(H))
            assert np.array_equal(d, Q.defects(tab, g["z"][w], N, simpson_sign=sign))
            if w % 2 == 1:      # perturbed windows sit off the knots, where derivatives are well defined
                for k in (0, 5, 11):
                    Jr = Q.hs_defect_jac(tab, xk[k], xn[k], u[k], simpson_sign=sign)
                    assert np.abs(J[k] - Jr).max() < 1e-12
                    Hr = Q.hs_lagrangian_hess(tab, xk[k], xn[k], u[k], lam[k], simpson_sign=sign)
                    assert np.abs(H[k] - Hr).max() < 1e-7 * max(1.0, np.abs(Hr).max())
                    assert np.array_equal(H[k], H[k].T) and not H[k][10:].any()
