"""Property-based tests (hypothesis) of the host-side logic: the table builder / interpolators of the C ABI against
the oracle port on random -- including non-monotone, to exercise the repair -- tables, the binary table cache, and the
batch sharding arithmetic.  No GPU needed: the table functions of libmpcb200.so are host code."""
import os
import tempfile

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import safe_autonomous_driving_mpc_b200 as M
from oracle import tracker_port as P


@st.composite
def tables(draw):
    K = draw(st.integers(min_value=3, max_value=40))
    # knot spacing may be zero or negative: trajectory_loader.py:27-30 repairs it to a strictly increasing sequence
    ds = draw(st.lists(st.floats(min_value=-0.05, max_value=2.0, allow_nan=False), min_size=K, max_size=K))
    cols = draw(st.lists(st.floats(min_value=-3.0, max_value=3.0, allow_nan=False), min_size=4 * K, max_size=4 * K))
    nu = draw(st.sampled_from([K - 1, K, K - 2]))
    us = draw(st.lists(st.floats(min_value=-5.0, max_value=5.0, allow_nan=False), min_size=2 * nu, max_size=2 * nu))
    X = np.empty((K, 5))
    X[:, 0] = np.cumsum(ds)
    X[:, 1:] = np.array(cols).reshape(K, 4)
    U = np.array(us).reshape(nu, 2)
    return X, U


@settings(max_examples=60, deadline=None)
@given(tab=tables(), probes=st.lists(st.floats(min_value=-0.5, max_value=1.5, allow_nan=False), min_size=1, max_size=12))
def test_table_interpolators_match_the_port(tab, probes):
    """mpcb_table_get_state / get_control == TrajectoryLoader.get_state / get_control of the reference as restated
    by the oracle (searchsorted-left segment choice, two-term formula, linear extrapolation below the first knot,
    last row / zero controls at and beyond s_max, controls on the first min(K, len(U)) knots)."""
    X, U = tab
    if U.shape[0] < 2:
        return
    L = M.TrajectoryLoader(X, U)
    R = P.RefTable(X, U)
    assert L.s_max == R.s_max
    span = R.s_max - R.s[0]
    for q in probes:
        s = R.s[0] + q * span                       # a bit before the first knot .. beyond the last
        np.testing.assert_allclose(L.get_state(s), R.get_state(s), rtol=0, atol=1e-12 * (1 + abs(s)))
        np.testing.assert_allclose(L.get_control(s), R.get_control(s), rtol=0, atol=1e-11)
    for s in (R.s[1], R.s[-2], R.s_max):            # exactly on knots
        np.testing.assert_allclose(L.get_state(float(s)), R.get_state(float(s)), rtol=0, atol=1e-12 * (1 + abs(s)))


@settings(max_examples=25, deadline=None)
@given(tab=tables())
def test_binary_table_cache_round_trip(tab):
    X, U = tab
    if U.shape[0] < 2:
        return
    L = M.TrajectoryLoader(X, U)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "t.mpct")
        L.save_binary(path)
        L2 = M.TrajectoryLoader(path)
    assert np.array_equal(L2.X_ref, np.ascontiguousarray(X)) and L2.s_max == L.s_max
    for s in np.linspace(X[0, 0] - 1.0, L.s_max + 1.0, 9):
        assert np.array_equal(L2.get_state(float(s)), L.get_state(float(s)))
        assert np.array_equal(L2.get_control(float(s)), L.get_control(float(s)))


@settings(max_examples=200, deadline=None)
@given(B=st.integers(min_value=0, max_value=200000), world=st.integers(min_value=1, max_value=16))
def test_shard_bounds_partition_the_batch(B, world):
    b = [M.sharding.shard_bounds(B, r, world) for r in range(world)]
    assert b[0][0] == 0 and b[-1][1] == B
    assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))           # contiguous, no gap, no overlap
    sizes = [hi - lo for lo, hi in b]
    assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1                # weak scaling: equal work per rank
    with pytest.raises(ValueError):
        M.sharding.shard_bounds(B, world, world)
