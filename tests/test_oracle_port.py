"""The CPU oracle (oracle/tracker_port.py) pinned against vectors produced by the unmodified reference
(tools/make_golden.py) and against the known-answer values of SURVEY.md Appendix B."""
import numpy as np
import pytest

from conftest import golden
from oracle import tracker_port as P


@pytest.mark.parametrize("i", [1, 2, 3])
def test_table_lookup_matches_reference(i, port_tables):
    z = golden(f"fn_traj{i}")
    tab = port_tables[i]
    gs = np.array([tab.get_state(s) for s in z["s_query"]])
    gc = np.array([tab.get_control(s) for s in z["s_query"]])
    # same expression, same operation order as scipy's interp1d -> bit-exact
    assert np.array_equal(gs, z["get_state"])
    assert np.array_equal(gc, z["get_control"])


@pytest.mark.parametrize("i", [1, 2, 3])
def test_model_functions_match_reference(i, port_tables):
    z = golden(f"fn_traj{i}")
    tab = port_tables[i]
    for t in range(len(z["x0"])):
        obs = [tuple(o) for o in z["obs_sv"][t, : z["n_obs"][t]]]
        assert np.array_equal(P.predict(tab, z["x0"][t], z["U"][t]), z["predict"][t])
        assert P.cost(tab, z["U"][t], z["x0"][t]) == z["cost"][t]
        c = P.constraint_values(tab, z["U"][t], z["x0"][t], obs)
        assert np.array_equal(c, z["constraints"][t, : len(c)])
        assert np.all(np.isnan(z["constraints"][t, len(c):]))


@pytest.mark.parametrize("name,i", [("solve_traj1", 1), ("solve_traj2", 2), ("solve_traj3", 3), ("solve_mc_traj3", 3)])
def test_warm_start_matches_reference(name, i, port_tables):
    z = golden(name)
    tab = port_tables[i]
    for t in range(len(z["x0"])):
        obs = [tuple(o) for o in z["obs_sv"][t, : z["n_obs"][t]]]
        assert np.array_equal(P.warm_start(tab, z["x0"][t], obs), z["U_init"][t])


def test_as_shipped_solve_matches_reference(port_tables):
    """solve() exactly as shipped (SLSQP ftol=1e-3, maxiter=15, FD gradients) reproduces the reference's
    iterate: same callables + same scipy => same floating-point trajectory."""
    z = golden("solve_traj1")
    tab = port_tables[1]
    for t in range(0, len(z["x0"]), 6):
        u0, pred, _sec, sol = P.solve_as_shipped(tab, z["x0"][t], [])
        np.testing.assert_allclose(sol.x, z["U_ship"][t], rtol=0, atol=1e-9)
        assert sol.status == z["ship_status"][t]
        np.testing.assert_allclose(pred, z["predX_ship"][t], rtol=0, atol=1e-9)


def test_appendix_b_known_answers(port_tables):
    """SURVEY.md Appendix B.1-B.3 (values printed by the unmodified reference)."""
    Ut = np.array([0.05, -1.0, -0.02, 0.5, 0.0, 0.0, 0.1, 1.5, -0.3, -2.0])
    t1, t2, t3 = port_tables[1], port_tables[2], port_tables[3]
    x = np.array([50.0, 0.1, 0.02, 0.01, 8.0])
    np.testing.assert_allclose(t1.get_state(50.0), [50.0, 0.0004124420896308335, 4.126146524266843e-05,
                                                    -0.0001772531162418013, 10.361183904781413], rtol=1e-15)
    np.testing.assert_allclose(t1.get_control(50.0), [0.00046633656814017144, 0.7578625114932006], rtol=1e-15)
    assert P.cost(t1, Ut, x) == pytest.approx(207.7498093959108, rel=1e-14)
    c = P.constraint_values(t1, Ut, x, [])
    np.testing.assert_allclose(c[:7], [0.268, 0.532, 0.2172029530196184, 0.5827970469803816, 0.16640590603923675,
                                       0.6335940939607633, 7.8], rtol=1e-13)
    np.testing.assert_allclose(P.predict(t1, x, Ut)[5], [57.96, 0.6359337681689039, 0.17641926916283268,
                                                          -0.023999999999999994, 7.799999999999999], rtol=1e-13)
    x = np.array([700.0, -0.05, 0.01, 0.0, 6.0])
    assert P.cost(t2, Ut, x) == pytest.approx(1592.0862662967418, rel=1e-14)
    c = P.constraint_values(t2, Ut, x, [(720.0, 4.0)])
    np.testing.assert_allclose(c[6::8], [10.89999999999991, 10.39000000000001, 10.010000000000014, 9.180000000000133,
                                         9.340000000000078], rtol=1e-12)
    x = np.array([1950.0, 0.0, 0.0, 0.0, 9.0])
    assert P.cost(t3, Ut, x) == pytest.approx(128.12470702267467, rel=1e-14)
    c = P.constraint_values(t3, Ut, x, [(1975.0, 3.0), (2000.0, 0.0)])
    np.testing.assert_allclose(c[36:45], [0.9042293754147012, -0.10422937541470112, 0.9886212694899416,
                                          -0.1886212694899415, 1.0730131635651818, -0.2730131635651818,
                                          5.84000000000019, 27.840000000000188, 8.8], rtol=1e-12)


def test_converged_oracle_reproduces_golden(port_tables):
    """The converged-oracle procedure of the port lands on the reference's converged answers (a few cases;
    the full sets are what the GPU tests compare against)."""
    z = golden("solve_mc_traj3")
    tab = port_tables[3]
    idx = [i for i in range(len(z["x0"])) if z["pinned"][i]][:4]
    for t in idx:
        obs = [tuple(o) for o in z["obs_sv"][t, : z["n_obs"][t]]]
        r = P.converged_oracle(tab, z["x0"][t], obs)
        assert r["pinned"]
        np.testing.assert_allclose(r["U"], z["U_conv"][t], atol=5e-5)
        assert r["J"] == pytest.approx(z["J_conv"][t], rel=1e-7, abs=1e-7)


@pytest.mark.parametrize("i,steps", [(1, 172), (2, 985), (3, 2294)])
def test_golden_closed_loop_shape(i, steps):
    z = golden(f"closed_loop_traj{i}")
    assert len(z["hist_u"]) == steps and len(z["hist_x"]) == steps + 1


def test_fsm_port_reproduces_reference_obstacle_log():
    """ObstacleFSMPort driven by the reference's own state history emits the reference's obstacle sets."""
    for i, cfg in ((2, P.FSM_TRAJ2), (3, P.FSM_TRAJ3)):
        z = golden(f"closed_loop_traj{i}")
        fsm = P.ObstacleFSMPort(True, True, **cfg)
        for t in range(len(z["hist_u"])):
            obs, tl = fsm.update(P.DT, z["hist_x"][t, 0], z["hist_x"][t, 4])
            assert len(obs) == z["n_obs"][t]
            for k, o in enumerate(obs):
                assert o["s"] == z["obs_sv"][t, k, 0] and o["v"] == z["obs_sv"][t, k, 1]
            assert tl == str(z["hist_tl"][t])


def test_monte_carlo_generator_is_deterministic(port_tables):
    a = P.monte_carlo_problems(port_tables[3], 65536)
    z = golden("solve_mc_traj3")
    m = len(z["n_obs"])
    assert m == 2048                                           # BASELINE.md 3 / SURVEY 8d: 2,048 problems with a converged answer
    assert np.array_equal(a[0][:m], z["x0"]) and np.array_equal(a[1][:m], z["obs_sv"])
    assert np.array_equal(a[2][:m], z["n_obs"])
    assert (z["pinned"] & (z["n_obs"] == 2)).sum() >= 200      # two-obstacle problems with a converged answer
    assert (np.arange(m) % 20 == 0).sum() == 103               # the full stress slice of the first 2,048
    assert (a[2] <= 2).all() and (a[2] >= 0).all()
