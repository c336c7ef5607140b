"""The reference's S2 calculus tests (dedalus/tests/test_sphere_calculus.py:24-400) restated against `dedalus_b200`, real dtype:
skew, convert-constant, average, gradient and Laplacian of spherical harmonics, MulCosine -- explicit, and implicit through LBVPs."""
import numpy as np
from scipy.special import sph_harm_y
import dedalus_b200 as d3

dtype = np.float64
radius = 1.37


def build_sphere(Nphi, Ntheta, dealias):
    c = d3.S2Coordinates('phi', 'theta')
    d = d3.Distributor(c, dtype=dtype)
    b = d3.SphereBasis(c, (Nphi, Ntheta), radius=radius, dealias=(dealias, dealias), dtype=dtype)
    phi, theta = d.local_grids(b, scales=dealias)
    return c, d, b, phi, theta


def check_explicit(dealias, Nphi=32, Ntheta=16):
    for layout in ('c', 'g'):                                            # test_skew_explicit
        c, d, b, phi, theta = build_sphere(Nphi, Ntheta, dealias)
        f = d.VectorField(c, bases=b)
        f.fill_random(layout='g')
        f.low_pass_filter(scales=0.75)
        f.change_layout(layout)
        g = d3.Skew(f).evaluate()
        assert np.allclose(g['g'][0], f['g'][1])
        assert np.allclose(g['g'][1], -f['g'][0])
    c, d, b, phi, theta = build_sphere(Nphi, Ntheta, dealias)
    f = d.Field()                                                        # test_convert_constant_scalar_explicit
    f['g'] = 1
    g = d3.Convert(f, b).evaluate()
    assert np.allclose(g['g'], 1)
    f = d.Field(bases=b)                                                 # test_sphere_average_scalar_explicit
    f.preset_scales(dealias)
    x = np.sin(theta)*np.cos(phi)
    z = np.cos(theta)
    f['g'] = 1 + x + z
    h = d3.Average(f, c).evaluate()
    assert np.allclose(h['g'], 1)
    m, l = 2, 2                                                          # test_gradient_scalar_explicit
    f = d.Field(bases=b)
    f.preset_scales(dealias)
    f['g'] = sph_harm_y(l, m, theta, phi).real
    u = d3.Gradient(f).evaluate()
    ug_phi = 1j*np.exp(2j*phi)*np.sqrt(15/(2*np.pi))*np.sin(theta)/2
    ug_theta = np.exp(2j*phi)*np.sqrt(15/(2*np.pi))*np.cos(theta)*np.sin(theta)/2
    ug = (np.array([ug_phi, ug_theta]) / radius).real
    u.change_scales(dealias)
    assert np.allclose(u['g'], ug)
    for rank in (0, 1, 2):                                               # test_cosine_explicit
        f = d.TensorField((c,)*rank, bases=b)
        f.fill_random(layout='g')
        f.low_pass_filter(scales=0.75)
        g = d3.MulCosine(f).evaluate()
        g.change_scales(dealias)
        f.change_scales(dealias)
        assert np.allclose(g['g'], np.cos(theta) * f['g'])
    m, l = 6, 10                                                         # test_laplacian_scalar_explicit
    f = d.Field(bases=b)
    f.preset_scales(dealias)
    f['g'] = sph_harm_y(l, m, theta, phi).real
    u = d3.Laplacian(f).evaluate()
    u.change_scales(dealias); f.change_scales(dealias)
    assert np.allclose(u['g'], -f['g']*(l*(l+1))/radius**2)


def check_implicit(dealias, Nphi=32, Ntheta=16):
    c, d, b, phi, theta = build_sphere(Nphi, Ntheta, dealias)
    f = d.VectorField(c, bases=b)                                        # test_skew_implicit
    f.fill_random(layout='g')
    f.low_pass_filter(scales=0.75)
    u = d.VectorField(c, bases=b)
    problem = d3.LBVP([u], namespace=locals())
    problem.add_equation("skew(u) = skew(f)")
    solver = problem.build_solver()
    solver.solve()
    u.change_scales(dealias)
    f.change_scales(dealias)
    assert np.allclose(u['g'], f['g'])
    for rank in (0, 1):                                                  # test_cosine_implicit
        f = d.TensorField((c,)*rank, bases=b)
        f.fill_random(layout='g')
        f.low_pass_filter(scales=0.75)
        u = d.TensorField((c,)*rank, bases=b)
        problem = d3.LBVP([u], namespace=dict(u=u, f=f, MulCosine=d3.MulCosine))
        problem.add_equation("u + MulCosine(u) = f + MulCosine(f)")
        solver = problem.build_solver()
        solver.solve()
        u.change_scales(dealias)
        f.change_scales(dealias)
        assert np.allclose(u['g'], f['g'])
    m, l = 5, 10                                                         # test_laplacian_scalar_implicit
    f = d.Field(bases=b)
    f.preset_scales(dealias)
    f['g'] = sph_harm_y(l, m, theta, phi).real
    u = d.Field(bases=b)
    tau = d.Field()
    problem = d3.LBVP([u, tau], namespace=locals())
    problem.add_equation("lap(u) + tau = f")
    problem.add_equation("ave(u) = 0")
    solver = problem.build_solver()
    solver.solve()
    u.change_scales(dealias); f.change_scales(dealias)
    assert np.allclose(u['g'], -f['g'] / (l*(l+1)) * radius**2)


def check_shell_gradient_scalar(dealias, Nphi=16, Ntheta=8, Nr=8):
    """test_spherical_calculus.py:43-57 with the shell basis (radii 0.5, 3): gradient of 3 x^2 + 2 y z in spherical components."""
    c = d3.SphericalCoordinates('phi', 'theta', 'r')
    d = d3.Distributor((c,), dtype=dtype)
    b = d3.ShellBasis(c, (Nphi, Ntheta, Nr), radii=(0.5, 3), dealias=(dealias, dealias, dealias), dtype=dtype)
    phi, theta, r = d.local_grids(b, scales=dealias)
    x, y, z = c.cartesian(phi, theta, r)
    f = d.Field(bases=b)
    f.preset_scales(dealias)
    f['g'] = 3*x**2 + 2*y*z
    u = d3.Gradient(f, c).evaluate()
    u.change_scales(dealias)
    ug = 0 * u['g']
    ug[2] = (6*x**2+4*y*z)/r
    ug[1] = -2*(y**3+x**2*(y-3*z)-y*z**2)/(r**2*np.sin(theta))
    ug[0] = 2*x*(-3*y+z)/(r*np.sin(theta))
    assert np.allclose(u['g'], ug)
