"""Sphere LBVP on the GPU: the balanced-height problem of the stock shallow-water script (gauge constant, average condition,
MulCosine on the right-hand side) followed by the IVP, against the unmodified reference (tests/golden/sphere_lbvp.npz)."""
import pytest
import sphere_cases as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["bal32", "bal64"])
def test_balanced_height_lbvp_then_ivp_matches_reference(golden, tag):
    S.check_balanced_shallow_water(golden("sphere_lbvp.npz"), tag)
