"""One-dimensional operator checks modelled on the reference's dedalus/tests/test_fourier_operators.py:24-100 and
test_jacobi_operators.py:25-160 (real dtype, analytic expectations): conversion of constants and between Jacobi bases (explicit and
through an LBVP), differentiation, interpolation, integration, averaging, lifting."""
import numpy as np
import dedalus_b200 as d3

REAL = np.float64


def _setup(make_basis):
    coord = d3.Coordinate('x')
    dist = d3.Distributor(coord, dtype=REAL)
    basis = make_basis(coord)
    return coord, dist, basis, dist.local_grid(basis, scale=1)


def _constant_converts(dist, basis):
    for layout in ('g', 'c'):
        one = dist.Field()
        one['g'] = 1
        one.change_layout(layout)
        assert np.allclose(d3.Convert(one, basis).evaluate()['g'], one['g'])


def check_fourier(N=10, bounds=(0.5, 1.666), dealias=1):
    """Fourier basis on an off-centre interval: f = 1 + sin(k x + 0.1) with k two periods per interval"""
    lo, hi = bounds
    k = 4 * np.pi / (hi - lo)
    coord, dist, basis, x = _setup(lambda c: d3.Fourier(c, size=N, bounds=bounds, dealias=dealias, dtype=REAL))
    _constant_converts(dist, basis)
    wave = dist.Field(bases=basis)
    wave['g'] = 1 + np.sin(k * x + 0.1)
    assert np.allclose(d3.Differentiate(wave, coord).evaluate()['g'], k * np.cos(k * x + 0.1))
    for point in (lo, hi, lo + (hi - lo) * np.random.rand()):
        assert np.allclose(d3.Interpolate(wave, coord, point).evaluate()['g'], 1 + np.sin(k * point + 0.1))
    assert np.allclose(d3.Integrate(wave, coord).evaluate()['g'], hi - lo)
    assert np.allclose(d3.Average(wave, coord).evaluate()['g'], 1)


def check_jacobi(N, a, b, k, dealias=1):
    """Jacobi basis (a + k, b + k) on the grid of (a, b): Chebyshev and Legendre families, k = 0, 1"""
    family = lambda bounds: (lambda c: d3.Jacobi(c, size=N, a0=a, b0=b, a=a + k, b=b + k, bounds=bounds, dealias=dealias))
    coord, dist, basis, x = _setup(family((0, 1)))
    _constant_converts(dist, basis)
    for dk in (0, 1, 2):
        higher = basis.derivative_basis(dk)
        for layout in ('g', 'c'):                         # explicit conversion upwards keeps the grid values
            smooth = dist.Field(bases=basis)
            smooth.fill_random(layout='g')
            smooth.low_pass_filter(scales=0.5)
            smooth.change_layout(layout)
            assert np.allclose(d3.Convert(smooth, higher).evaluate()['g'], smooth['g'])
        data = dist.Field(bases=higher)                   # implicit conversion downwards: solve  g = data  for g in the lower basis
        data.fill_random(layout='g')
        data.low_pass_filter(scales=0.5)
        lower = dist.Field(bases=basis)
        bvp = d3.LBVP([lower], namespace=dict(g=lower, f=data))
        bvp.add_equation("g = f")
        bvp.build_solver().solve()
        assert np.allclose(lower['g'], data['g'])
    quintic = dist.Field(bases=basis)
    quintic['g'] = x ** 5
    assert np.allclose(d3.Differentiate(quintic, coord).evaluate()['g'], 5 * x ** 4)
    for point in (0, 1, np.random.rand()):
        assert np.allclose(d3.Interpolate(quintic, coord, point).evaluate()['g'], point ** 5)
    coord, dist, basis, x = _setup(family((0, 3)))
    sextic_slope = dist.Field(bases=basis)
    sextic_slope['g'] = 6 * x ** 5
    assert np.allclose(d3.Integrate(sextic_slope, coord).evaluate()['g'], 3 ** 6)
    assert np.allclose(d3.Average(sextic_slope, coord).evaluate()['g'], 3 ** 6 / 3)
    target_basis = basis.derivative_basis(k)
    for mode in (-1, -2):                                 # lifting a constant into one mode of the target basis
        expected = dist.Field(bases=target_basis)
        expected['c'][mode] = 2
        amplitude = dist.Field()
        amplitude['g'] = 2
        assert np.allclose(d3.Lift(amplitude, target_basis, mode).evaluate()['g'], expected['g'])
