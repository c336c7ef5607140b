"""The reference's 1-D operator tests restated against `dedalus_b200`, real dtype, analytic expectations as in the reference:
dedalus/tests/test_fourier_operators.py:24-100 (convert constant, differentiate, interpolate, integrate, average) and
dedalus/tests/test_jacobi_operators.py:25-160 (the same plus explicit / implicit basis conversion and lift)."""
import numpy as np
import dedalus_b200 as d3

dtype = np.float64


def build_fourier(N, bounds, dealias):
    c = d3.Coordinate('x')
    d = d3.Distributor(c, dtype=dtype)
    b = d3.Fourier(c, size=N, bounds=bounds, dealias=dealias, dtype=dtype)
    x = d.local_grid(b, scale=1)
    return c, d, b, x


def check_fourier(N=10, bounds=(0.5, 1.666), dealias=1):
    k = 4 * np.pi / (bounds[1] - bounds[0])
    for layout in ('g', 'c'):                                           # test_fourier_convert_constant
        c, d, b, x = build_fourier(N, bounds, dealias)
        f = d.Field()
        f['g'] = 1
        f.change_layout(layout)
        g = d3.Convert(f, b).evaluate()
        assert np.allclose(g['g'], f['g'])
    c, d, b, x = build_fourier(N, bounds, dealias)
    f = d.Field(bases=b)
    f['g'] = 1 + np.sin(k*x+0.1)
    g = d3.Differentiate(f, c).evaluate()                               # test_fourier_differentiate
    assert np.allclose(g['g'], k*np.cos(k*x+0.1))
    for p in [bounds[0], bounds[1], bounds[0] + (bounds[1] - bounds[0]) * np.random.rand()]:      # test_fourier_interpolate
        g = d3.Interpolate(f, c, p).evaluate()
        assert np.allclose(g['g'], 1 + np.sin(k*p+0.1))
    g = d3.Integrate(f, c).evaluate()                                   # test_fourier_integrate
    assert np.allclose(g['g'], bounds[1] - bounds[0])
    g = d3.Average(f, c).evaluate()                                     # test_fourier_average
    assert np.allclose(g['g'], 1)


def build_jacobi(N, a, b, k, bounds, dealias):
    c = d3.Coordinate('x')
    d = d3.Distributor(c, dtype=dtype)
    b = d3.Jacobi(c, size=N, a0=a, b0=b, a=a+k, b=b+k, bounds=bounds, dealias=dealias)
    x = d.local_grid(b, scale=1)
    return c, d, b, x


def check_jacobi(N, a, b, k, dealias=1):
    for layout in ('g', 'c'):                                           # test_jacobi_convert_constant
        c, d, bs, x = build_jacobi(N, a, b, k, (0, 1), dealias)
        f = d.Field()
        f['g'] = 1
        f.change_layout(layout)
        g = d3.Convert(f, bs).evaluate()
        assert np.allclose(g['g'], f['g'])
    for dk in (0, 1, 2):
        for layout in ('g', 'c'):                                       # test_jacobi_convert
            c, d, bs, x = build_jacobi(N, a, b, k, (0, 1), dealias)
            f = d.Field(bases=bs)
            f.fill_random(layout='g')
            f.low_pass_filter(scales=0.5)
            f.change_layout(layout)
            g = d3.Convert(f, bs.derivative_basis(dk)).evaluate()
            assert np.allclose(g['g'], f['g'])
        c, d, bs, x = build_jacobi(N, a, b, k, (0, 1), dealias)         # test_jacobi_convert_implicit
        f = d.Field(bases=bs.derivative_basis(dk))
        f.fill_random(layout='g')
        f.low_pass_filter(scales=0.5)
        g = d.Field(bases=bs)
        problem = d3.LBVP([g], namespace=locals())
        problem.add_equation("g = f")
        solver = problem.build_solver()
        solver.solve()
        assert np.allclose(g['g'], f['g'])
    c, d, bs, x = build_jacobi(N, a, b, k, (0, 1), dealias)
    f = d.Field(bases=bs)
    f['g'] = x**5
    g = d3.Differentiate(f, c).evaluate()                               # test_jacobi_differentiate
    assert np.allclose(g['g'], 5*x**4)
    for p in [0, 1, np.random.rand()]:                                  # test_jacobi_interpolate
        fp = d3.Interpolate(f, c, p).evaluate()
        assert np.allclose(fp['g'], p**5)
    c, d, bs, x = build_jacobi(N, a, b, k, (0, 3), dealias)
    f = d.Field(bases=bs)
    f['g'] = 6 * x**5
    assert np.allclose(d3.Integrate(f, c).evaluate()['g'], 3**6)        # test_jacobi_integrate
    assert np.allclose(d3.Average(f, c).evaluate()['g'], 3**6 / 3)      # test_jacobi_average
    for n in (-1, -2):                                                  # test_jacobi_lift
        c, d, bs, x = build_jacobi(N, a, b, k, (0, 3), dealias)
        lift_basis = bs.derivative_basis(k)
        f = d.Field(bases=lift_basis)
        f['c'][n] = 2
        tau = d.Field()
        tau['g'] = 2
        g = d3.Lift(tau, lift_basis, n).evaluate()
        assert np.allclose(g['g'], f['g'])
