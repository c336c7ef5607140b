"""The reference's own Cartesian LBVP tests (tests/lbvp_cases.py) and ufunc tests (tests/grid_operator_cases.py) through the CPU emulation."""
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
