"""The reference's Matsolver / Transpose plugin contracts over the device kernels, through the CPU emulation."""
import pytest
from emu import emu_lib as E
import plugin_cases as P


@pytest.fixture(autouse=True)
def emulation():
    E.install()
    yield
    E.uninstall()


def test_matsolver_registry_contract():
    P.check_matsolvers()


def test_transpose_plugin_single_rank():
    P.check_transpose_single_rank()
