"""Spherical shell on the GPU: field transforms (T6: radial Jacobi transform + regularity recombination on top of the sphere
chain) against reference vectors (tests/golden/shell.npz), the dense per-l pencil kernels, and the shell-convection IVP of
BASELINE config 5 against reference states (tests/golden/shell_ivp.npz)."""
import pytest
import shell_cases as SC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_shell_field_transforms(golden, tag):
    SC.check_shell_field_transforms(golden("shell.npz"), tag)


def test_dense_kernels_against_numpy():
    from test_gpu_2_sphere import _CudaArrays
    SC.check_dense_kernels(_CudaArrays())


def test_shell_pencil_matrices(golden):
    SC.check_shell_pencil_matrices(golden("shell_ivp.npz"))


@pytest.mark.parametrize("tag,scheme", [("a_sbdf2", "SBDF2"), ("a_rk222", "RK222"), ("b_sbdf2", "SBDF2")])
def test_shell_convection_matches_reference(golden, tag, scheme):
    """BASELINE config 5's problem (shell convection) at 16 x 8 x 6 and 32 x 16 x 12 against the reference."""
    solver = SC.check_shell_convection(golden("shell_ivp.npz"), tag, scheme)
    assert solver.bset.last_verify < 1e-12
