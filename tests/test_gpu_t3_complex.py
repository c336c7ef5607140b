"""Complex-dtype pencil path (T3) on the GPU: the reference's own complex IVP test (tests/test_ivp.py:18-49, every scheme) and
a ComplexFourier x ChebyshevT problem with complex coefficients against reference states (tests/golden/complex_cgl.npz)."""
import numpy as np, pytest
import dedalus_b200 as d3
import complex_cases as CC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("timestepper", list(d3.schemes.keys()))
def test_heat_periodic_complex_every_timestepper(timestepper):
    CC.check_heat_periodic(timestepper)


@pytest.mark.parametrize("tag,scheme", [("rk222", "RK222"), ("sbdf2", "SBDF2")])
def test_complex_ginzburg_landau_matches_reference(golden, tag, scheme):
    solver = CC.check_ginzburg_landau(golden("complex_cgl.npz"), tag, scheme)
    assert solver.bset.last_verify < 1e-10
