"""Spherical-shell field transforms (T6: radial Jacobi transform + regularity recombination on top of the sphere chain) on the
GPU against reference vectors (tests/golden/shell.npz)."""
import pytest
import shell_cases as SC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_shell_field_transforms(golden, tag):
    SC.check_shell_field_transforms(golden("shell.npz"), tag)
