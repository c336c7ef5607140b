"""The reference's Cartesian LBVP tests (tests/lbvp_cases.py) through the CPU emulation of the kernels."""
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
