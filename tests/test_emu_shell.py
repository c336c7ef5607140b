"""Spherical-shell field transforms (T6) through the CPU emulation of the kernels, against reference vectors
(tests/golden/shell.npz); GPU version: tests/test_gpu_t6_shell.py."""
import pytest
from emu import emu_lib as E
import shell_cases as SC


@pytest.fixture(autouse=True)
def emulation():
    E.install()
    yield
    E.uninstall()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_shell_field_transforms(golden, tag):
    SC.check_shell_field_transforms(golden("shell.npz"), tag)


def test_intertwiner_orthogonal():
    SC.check_intertwiner_orthogonal()
