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


def test_reference_operator_cases():
    """The reference's Fourier / Jacobi operator tests (test_fourier_operators.py, test_jacobi_operators.py)."""
    import operator_cases as O
    O.check_fourier()
    for N in (8, 9):
        for a, b in ((-1/2, -1/2), (0, 0)):
            for k in (0, 1):
                O.check_jacobi(N, a, b, k)
