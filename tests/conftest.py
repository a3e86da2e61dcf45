import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def traj_path(i):
    return os.path.join(DATA, f"trajectory{i}.npz")


@pytest.fixture(scope="session")
def port_tables():
    from oracle import tracker_port as P
    return {i: P.RefTable.from_npz(traj_path(i)) for i in (1, 2, 3)}


@pytest.fixture(scope="session")
def gpu_trackers():
    """One BatchedTracker per trajectory.  No skip-on-failure: a missing library or device is an error."""
    import safe_autonomous_driving_mpc_b200 as M
    out = {}
    for i in (1, 2, 3):
        L = M.TrajectoryLoader(traj_path(i))
        out[i] = (L, M.BatchedTracker(L))
    return out


# Execution shapes of the solve path.  "default": what a caller gets (batches up to coop_max_batch run one warp per
# problem); "bulk": the thread-per-problem first pass that carries the benchmark, forced for every batch size, with the
# warp-per-problem robust pass; "bulk_thread2": both passes thread-per-problem.
VARIANTS = {"default": {}, "bulk": dict(coop_max_batch=0), "bulk_thread2": dict(coop_max_batch=0, coop_pass2=0)}


@pytest.fixture(scope="session")
def tracker_variants(gpu_trackers):
    import safe_autonomous_driving_mpc_b200 as M
    out = {"default": gpu_trackers}
    for name, kw in VARIANTS.items():
        if name == "default":
            continue
        out[name] = {i: (gpu_trackers[i][0], M.BatchedTracker(gpu_trackers[i][0], **kw)) for i in (1, 2, 3)}
    return out
