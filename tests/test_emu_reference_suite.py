"""Cases of the reference's OWN test suite restated against dedalus_b200 (lbvp_cases, grid_operator_cases, operator_cases), through the CPU emulation."""
import pytest
from emu import emu_lib as E
import lbvp_cases as L


@pytest.fixture(autouse=True)
def emulation():
    E.install()
    yield
    E.uninstall()


def test_algebraic():
    L.check_algebraic()


def test_poisson_fourier():
    L.check_poisson_fourier()


@pytest.mark.parametrize("a,b", [(-1/2, -1/2), (0, 0)])
def test_poisson_jacobi(a, b):
    L.check_poisson_jacobi(a, b)


@pytest.mark.parametrize("a,b", [(-1/2, -1/2), (0, 0)])
def test_jacobi_ufunc_field(a, b):
    import grid_operator_cases as G
    G.check_jacobi_ufunc_field(a, b)


def test_shell_ufunc_field_and_operator():
    import grid_operator_cases as G
    G.check_shell_ufuncs()


def test_fourier_operators():
    import operator_cases as O
    O.check_fourier()


@pytest.mark.parametrize("N", [8, 9])
@pytest.mark.parametrize("a,b", [(-1/2, -1/2), (0, 0)])
@pytest.mark.parametrize("k", [0, 1])
def test_jacobi_operators(N, a, b, k):
    import operator_cases as O
    O.check_jacobi(N, a, b, k)


@pytest.mark.parametrize("dealias", [1, 3/2])
def test_cfl_operators(dealias):
    import cfl_cases as Cc
    Cc.check_cfl_1d('fourier', dealias)
    Cc.check_cfl_1d('chebyshev', dealias)
    Cc.check_cfl_fourier_chebyshev(dealias)
    Cc.check_cfl_sphere(dealias)
    Cc.check_cfl_shell(dealias)


@pytest.mark.parametrize("dealias", [1, 3/2])
@pytest.mark.parametrize("safety", [0.2, 0.4])
def test_full_cfl_fourier_chebyshev(dealias, safety):
    import cfl_cases as Cc
    Cc.check_full_cfl_fourier_chebyshev(dealias, safety)


@pytest.mark.parametrize("kind", ["FF", "FC"])
def test_cartesian_skew(kind):
    import cartesian_operator_cases as K
    K.check_skew(kind)


@pytest.mark.parametrize("kind", ["FF", "FC", "FFF", "FFC"])
def test_cartesian_trace_and_transpose(kind):
    import cartesian_operator_cases as K
    K.check_trace_and_transpose(kind)


def test_cartesian_curls():
    import cartesian_operator_cases as K
    K.check_curls()


@pytest.mark.parametrize("dealias", [1, 3/2])
def test_sphere_calculus_explicit(dealias):
    import sphere_calculus_cases as S
    S.check_explicit(dealias)


@pytest.mark.parametrize("dealias", [1, 3/2])
def test_sphere_calculus_implicit(dealias):
    import sphere_calculus_cases as S
    S.check_implicit(dealias)


@pytest.mark.parametrize("dealias", [1, 3/2])
def test_shell_gradient_scalar(dealias):
    import sphere_calculus_cases as S
    S.check_shell_gradient_scalar(dealias)


@pytest.mark.parametrize("dealias", [1, 3/2])
def test_shell_calculus(dealias):
    import sphere_calculus_cases as S
    S.check_shell_calculus(dealias)


@pytest.mark.parametrize("k", [0, 1])
@pytest.mark.parametrize("dealias", [1, 3/2])
def test_shell_operators(k, dealias):
    import sphere_calculus_cases as S
    S.check_shell_operators(k, dealias)


@pytest.mark.parametrize("a0,b0", [(-1/2, -1/2), (0, 0)])
@pytest.mark.parametrize("k_ncc", [0, 1])
def test_solve_jacobi_ncc(a0, b0, k_ncc):
    L.check_solve_jacobi_ncc(a0, b0, k_ncc)


@pytest.mark.parametrize("k", [0, 1])
def test_shell_implicit_trace_and_transpose(k):
    import sphere_calculus_cases as S
    S.check_shell_implicit(k, 3/2)


def test_shell_arithmetic():
    """At the reference's sizes (8, 10, 6): Ntheta > Nphi / 2, the folded packing with shift > 0."""
    import sphere_calculus_cases as S
    S.check_shell_arithmetic(3/2, Nphi=8, Ntheta=10, Nr=6)
