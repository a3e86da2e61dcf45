"""The numpy model of the GPU algorithm (oracle/sqp_admm_model.py) against the golden converged answers: guards the
algorithm's constants on the CPU so that a GPU failure can be split into 'algorithm' vs 'kernel'."""
import numpy as np
import pytest

from conftest import golden
from oracle import sqp_admm_model as A


@pytest.mark.parametrize("i", [1, 3])
def test_assembly_matches_reference_functions(i, port_tables):
    z = golden(f"fn_traj{i}")
    asm = A.assemble(port_tables[i], z["x0"], z["U"])
    np.testing.assert_allclose(asm["X"], z["predict"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(asm["cost"], z["cost"], rtol=1e-13)


def test_residual_jacobian_against_central_differences(port_tables):
    z = golden("fn_traj2")
    tab = port_tables[2]
    x0, U = z["x0"][:32], z["U"][:32]
    asm = A.assemble(tab, x0, U)
    eps = 1e-6
    for k in range(10):
        Up, Um = U.copy(), U.copy()
        Up[:, k] += eps
        Um[:, k] -= eps
        fd = (A.assemble(tab, x0, Up)["r"] - A.assemble(tab, x0, Um)["r"]) / (2 * eps)
        np.testing.assert_allclose(asm["Jr"][:, :, k], fd, atol=2e-6)


def test_u1_last_step_is_zero_at_optimum():
    """u1_4 moves only k_5, which appears nowhere -> exactly 0 at the optimum."""
    z = golden("solve_mc_traj3")
    assert np.max(np.abs(z["U_conv"][z["pinned"], 8])) < 5e-5
