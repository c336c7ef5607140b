"""The reference's Matsolver / Transpose plugin contracts over the device kernels, on the GPU."""
import pytest
import plugin_cases as P

pytestmark = pytest.mark.gpu


def test_matsolver_registry_contract():
    P.check_matsolvers()


def test_transpose_plugin_single_rank():
    P.check_transpose_single_rank()


def test_reference_lbvp_cases():
    """The reference's own Cartesian LBVP tests (dedalus/tests/test_lbvp.py:38-111)."""
    import lbvp_cases as L
    L.check_algebraic()
    L.check_poisson_fourier()
    L.check_poisson_jacobi(-1/2, -1/2)
    L.check_poisson_jacobi(0, 0)


def test_reference_ufunc_cases():
    """The reference's ufunc tests (dedalus/tests/test_grid_operators.py:35-85), Jacobi and shell, N as in the reference."""
    import grid_operator_cases as G
    G.check_jacobi_ufunc_field(-1/2, -1/2)
    G.check_jacobi_ufunc_field(0, 0)
    G.check_shell_ufuncs(N=16)
