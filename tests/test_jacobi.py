"""Host Jacobi algebra vs arrays dumped from the reference (tools/jacobi.py:203-260)."""
import numpy as np, pytest
from dedalus_b200 import jacobi


@pytest.mark.parametrize("ab", [(-0.5, -0.5), (0.5, 0.5), (1.5, 1.5), (0.0, 0.0)])
def test_operators_match_reference(golden, ab):
    g = golden("transforms.npz"); a, b = ab; N = 12
    tol = dict(rtol=1e-13, atol=1e-13)
    assert np.allclose(jacobi.differentiation_matrix(N, a, b).toarray(), g[f"jop_D_{a}_{b}"], **tol)
    assert np.allclose(jacobi.conversion_matrix(N, a, b, a + 1, b + 1).toarray(), g[f"jop_C_{a}_{b}"], **tol)
    assert np.allclose(jacobi.jacobi_matrix(N, a, b).toarray(), g[f"jop_Z_{a}_{b}"], **tol)
    assert np.allclose(jacobi.integration_vector(N, a, b), g[f"jop_int_{a}_{b}"], rtol=1e-12, atol=1e-13)
    z, w = jacobi.gauss_grid(N, a, b)
    assert np.allclose(z, g[f"jop_grid_{a}_{b}"], **tol)
    assert np.allclose(w, g[f"jop_wts_{a}_{b}"], rtol=1e-12, atol=1e-13)
    assert np.allclose(jacobi.polynomials(N, a, b, np.array([-1.0, 0.3, 1.0])), g[f"jop_pm1_{a}_{b}"], rtol=1e-12, atol=1e-12)
