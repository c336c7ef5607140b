"""The reference's Matsolver / Transpose plugin contracts over the device kernels, on the GPU."""
import pytest
import plugin_cases as P

pytestmark = pytest.mark.gpu


def test_matsolver_registry_contract():
    P.check_matsolvers()


def test_transpose_plugin_single_rank():
    P.check_transpose_single_rank()
