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
    for a0, b0 in ((-1/2, -1/2), (0, 0)):
        for k_ncc in (0, 1):
            L.check_solve_jacobi_ncc(a0, b0, k_ncc)            # test_cartesian_ncc.py:115-135


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


def test_reference_cfl_cases():
    """The reference's CFL tests (dedalus/tests/test_cfl.py), real dtype: operator on five kinds of bases and the tool end to end."""
    import cfl_cases as Cc
    for dealias in (1, 3/2):
        Cc.check_cfl_1d('fourier', dealias)
        Cc.check_cfl_1d('chebyshev', dealias)
        Cc.check_cfl_fourier_chebyshev(dealias)
        Cc.check_cfl_sphere(dealias)
        Cc.check_cfl_shell(dealias, N=16)
        for safety in (0.2, 0.4):
            Cc.check_full_cfl_fourier_chebyshev(dealias, safety)


def test_reference_cartesian_tensor_operator_cases():
    """The reference's skew / trace / transpose tests (dedalus/tests/test_cartesian_operators.py:93-250), explicit and implicit."""
    import cartesian_operator_cases as K
    for kind in ("FF", "FC"):
        K.check_skew(kind)
    for kind in ("FF", "FC", "FFF", "FFC"):
        K.check_trace_and_transpose(kind)
    K.check_curls()


def test_reference_sphere_calculus_cases():
    """The reference's S2 calculus tests (dedalus/tests/test_sphere_calculus.py), explicit and through LBVPs."""
    import sphere_calculus_cases as S
    for dealias in (1, 3/2):
        S.check_explicit(dealias)
        S.check_implicit(dealias)
        S.check_shell_gradient_scalar(dealias)
        S.check_shell_calculus(dealias)
        for k in (0, 1):
            S.check_shell_operators(k, dealias)
        for k in (0, 1):
            S.check_shell_implicit(k, dealias)
        S.check_shell_arithmetic(dealias, Nphi=8, Ntheta=10, Nr=6)
