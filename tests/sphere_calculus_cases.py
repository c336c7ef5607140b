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


def _shell(dealias, Nphi=16, Ntheta=8, Nr=8):
    c = d3.SphericalCoordinates('phi', 'theta', 'r')
    d = d3.Distributor((c,), dtype=dtype)
    b = d3.ShellBasis(c, (Nphi, Ntheta, Nr), radii=(0.5, 3), dealias=(dealias, dealias, dealias), dtype=dtype)
    phi, theta, r = d.local_grids(b, scales=dealias)
    x, y, z = c.cartesian(phi, theta, r)
    return c, d, b, phi, theta, r, x, y, z


def check_shell_calculus(dealias):
    """test_spherical_calculus.py:78-96, 120-131, 175-186, 206-223 with the shell basis: gradient of a gradient, divergence of a
    gradient, Laplacians of a scalar and of a vector field, against the analytic answers of the reference's tests."""
    c, d, b, phi, theta, r, x, y, z = _shell(dealias)
    grad = lambda A: d3.Gradient(A, c)
    f = d.Field(bases=b)
    f.preset_scales(dealias)
    f['g'] = 3*x**2 + 2*y*z
    T = grad(grad(f)).evaluate()                                         # test_gradient_vector
    T.change_scales(dealias)
    Tg = 0 * T['g']
    Tg[2,2] = (6*x**2+4*y*z)/r**2
    Tg[2,1] = Tg[1,2] = -2*(y**3+x**2*(y-3*z)-y*z**2)/(r**3*np.sin(theta))
    Tg[2,0] = Tg[0,2] = 2*x*(z-3*y)/(r**2*np.sin(theta))
    Tg[1,1] = 6*x**2/(r**2*np.sin(theta)**2) - (6*x**2+4*y*z)/r**2
    Tg[1,0] = Tg[0,1] = -2*x*(x**2+y**2+3*y*z)/(r**3*np.sin(theta)**2)
    Tg[0,0] = 6*y**2/(x**2+y**2)
    assert np.allclose(T['g'], Tg)
    f = d.Field(bases=b)                                                 # test_divergence_vector
    f.preset_scales(dealias)
    f['g'] = x**3 + 2*y**3 + 3*z**3
    h = d3.Divergence(grad(f)).evaluate()
    h.change_scales(dealias)
    assert np.allclose(h['g'], 6*x + 12*y + 18*z)
    f = d.Field(bases=b)                                                 # test_laplacian_scalar
    f.preset_scales(dealias)
    f['g'] = x**4 + 2*y**4 + 3*z**4
    h = d3.Laplacian(f, c).evaluate()
    h.change_scales(dealias)
    assert np.allclose(h['g'], 12*x**2+24*y**2+36*z**2)
    u = d.VectorField(c, bases=b)                                        # test_laplacian_vector
    u.preset_scales(dealias)
    ct, st, cp, sp = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)
    u['g'][2] = r**2*st*(2*ct**2*cp-r*ct**3*sp+r**3*cp**3*st**5*sp**3+r*ct*st**2*(cp**3+sp**3))
    u['g'][1] = r**2*(2*ct**3*cp-r*cp**3*st**4+r**3*ct*cp**3*st**5*sp**3-1/16*r*np.sin(2*theta)**2*(-7*sp+np.sin(3*phi)))
    u['g'][0] = r**2*sp*(-2*ct**2+r*ct*cp*st**2*sp-r**3*cp**2*st**5*sp**3)
    v = d3.Laplacian(u, c).evaluate()
    v.change_scales(dealias)
    vg = 0 * v['g']
    vg[2] = 2*(2+3*r*ct)*cp*st+1/2*r**3*st**4*(4*np.sin(2*phi)+np.sin(4*phi))
    vg[1] = 2*r*(-3*cp*st**2+sp)+1/2*ct*(8*cp+r**3*st**3*(4*np.sin(2*phi)+np.sin(4*phi)))
    vg[0] = 2*r*ct*cp+2*sp*(-2-r**3*(2+np.cos(2*phi))*st**3*sp)
    assert np.allclose(v['g'], vg)


def check_shell_operators(k, dealias, Nphi=8, Ntheta=4, Nr=10):
    """test_spherical_operators.py:120-262 with the shell basis (radii 0.5, 1.5; fields in the radial basis k): sums of fields in
    different radial bases (conversion), trace and transpose of a gradient, in both layouts."""
    c = d3.SphericalCoordinates('phi', 'theta', 'r')
    d = d3.Distributor((c,), dtype=dtype)
    b = d3.ShellBasis(c, (Nphi, Ntheta, Nr), radii=(0.5, 1.5), k=k, dealias=(dealias, dealias, dealias), dtype=dtype)
    phi, theta, r = d.local_grids(b, scales=dealias)
    x, y, z = c.cartesian(phi, theta, r)
    ct, st, cp, sp = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)

    def vector():
        u = d.VectorField(c, bases=b)
        u.preset_scales(dealias)
        u['g'][2] = r**2*st*(2*ct**2*cp-r*ct**3*sp+r**3*cp**3*st**5*sp**3+r*ct*st**2*(cp**3+sp**3))
        u['g'][1] = r**2*(2*ct**3*cp-r*cp**3*st**4+r**3*ct*cp**3*st**5*sp**3-1/16*r*np.sin(2*theta)**2*(-7*sp+np.sin(3*phi)))
        u['g'][0] = r**2*sp*(-2*ct**2+r*ct*cp*st**2*sp-r**3*cp**2*st**5*sp**3)
        return u
    for layout in ('c', 'g'):
        f = d.Field(bases=b)                                             # test_convert_scalar
        f.preset_scales(dealias)
        f['g'] = 3*x**2 + 2*y*z
        g = d3.Laplacian(f, c).evaluate()
        g.change_scales(dealias)
        f.change_layout(layout); g.change_layout(layout)
        h = (f + g).evaluate()
        h.change_scales(dealias); f.change_scales(dealias); g.change_scales(dealias)
        assert np.allclose(h['g'], f['g'] + g['g'])
        u = vector()                                                     # test_convert_vector
        v = d3.Laplacian(u, c).evaluate()
        v.change_scales(dealias)
        u.change_layout(layout); v.change_layout(layout)
        w = (u + v).evaluate()
        w.change_scales(dealias); u.change_scales(dealias); v.change_scales(dealias)
        assert np.allclose(w['g'], u['g'] + v['g'])
        u = vector()                                                     # test_explicit_trace_tensor
        T = d3.Gradient(u, c).evaluate()
        T.change_scales(dealias)
        fg = T['g'][0,0] + T['g'][1,1] + T['g'][2,2]
        T.change_layout(layout)
        f = d3.Trace(T).evaluate()
        f.change_scales(dealias)
        assert np.allclose(f['g'], fg)
        T = d3.Gradient(vector(), c).evaluate()                          # test_explicit_transpose_tensor
        T.change_scales(dealias)
        Tg = np.transpose(np.copy(T['g']), (1,0,2,3,4))
        T.change_layout(layout)
        T = d3.TransposeComponents(T).evaluate()
        T.change_scales(dealias)
        assert np.allclose(T['g'], Tg)


def check_shell_implicit(k, dealias, Nphi=8, Ntheta=4, Nr=10):
    """test_spherical_operators.py:191-208 and 262-283 with the shell basis: LBVPs trace(I*f) = 3 g with a radial-basis identity
    tensor, and transpose(Tt) = T for the gradient of a vector field."""
    c = d3.SphericalCoordinates('phi', 'theta', 'r')
    d = d3.Distributor((c,), dtype=dtype)
    b = d3.ShellBasis(c, (Nphi, Ntheta, Nr), radii=(0.5, 1.5), k=k, dealias=(dealias, dealias, dealias), dtype=dtype)
    phi, theta, r = d.local_grids(b, scales=dealias)
    x, y, z = c.cartesian(phi, theta, r)
    f = d.Field(bases=b)
    g = d.Field(bases=b)
    g.preset_scales(dealias)
    g['g'] = 3*x**2 + 2*y*z
    I = d.TensorField((c, c), bases=b.radial_basis)
    I['g'][0,0] = I['g'][1,1] = I['g'][2,2] = 1
    problem = d3.LBVP([f])
    problem.add_equation((d3.Trace(I*f), 3*g))
    solver = problem.build_solver()
    solver.solve()
    assert np.allclose(f['c'], g['c'])
    ct, st, cp, sp = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)
    u = d.VectorField(c, bases=b)
    u.preset_scales(dealias)
    u['g'][2] = r**2*st*(2*ct**2*cp-r*ct**3*sp+r**3*cp**3*st**5*sp**3+r*ct*st**2*(cp**3+sp**3))
    u['g'][1] = r**2*(2*ct**3*cp-r*cp**3*st**4+r**3*ct*cp**3*st**5*sp**3-1/16*r*np.sin(2*theta)**2*(-7*sp+np.sin(3*phi)))
    u['g'][0] = r**2*sp*(-2*ct**2+r*ct*cp*st**2*sp-r**3*cp**2*st**5*sp**3)
    T = d3.Gradient(u, c).evaluate()
    T.change_scales(dealias)
    Ttg = np.transpose(np.copy(T['g']), (1,0,2,3,4))
    Tt = d.TensorField((c, c), bases=T.unique_bases())
    problem = d3.LBVP([Tt])
    problem.add_equation((d3.TransposeComponents(Tt), T))
    solver = problem.build_solver()
    solver.solve()
    Tt.change_scales(dealias)
    assert np.allclose(Tt['g'], Ttg)


def check_shell_arithmetic(dealias, Nphi=16, Ntheta=8, Nr=8):
    """test_spherical_arithmetic.py:111-250 with the shell basis, real dtype: dot products (vector.vector, tensor.vector), products
    with numbers and tensor products up to rank 4, against numpy on the grid values."""
    c = d3.SphericalCoordinates('phi', 'theta', 'r')
    d = d3.Distributor((c,), dtype=dtype)
    b = d3.ShellBasis(c, (Nphi, Ntheta, Nr), radii=(0.5, 3), dealias=(dealias, dealias, dealias), dtype=dtype)
    phi, theta, r = d.local_grids(b, scales=dealias)
    x, y, z = c.cartesian(phi, theta, r)

    def at(f):
        f.change_scales(dealias)
        return np.array(f['g'])
    f = d.Field(bases=b); f.preset_scales(dealias); f['g'] = z
    ez = d3.Gradient(f, c).evaluate()
    u = d.VectorField(c, bases=b); u.preset_scales(dealias)
    u['g'][2] = (6*x**2+4*y*z)/r
    u['g'][1] = -2*(y**3+x**2*(y-3*z)-y*z**2)/(r**2*np.sin(theta))
    u['g'][0] = 2*x*(-3*y+z)/(r*np.sin(theta))
    h = d3.DotProduct(ez, u).evaluate()                                  # test_dot_product_vector_vector
    assert np.allclose(at(h), np.sum(at(ez)*at(u), axis=0))
    T = d.TensorField((c, c), bases=b); T.preset_scales(dealias)
    T['g'][2,2] = (6*x**2+4*y*z)/r**2
    T['g'][2,1] = T['g'][1,2] = -2*(y**3+x**2*(y-3*z)-y*z**2)/(r**3*np.sin(theta))
    T['g'][2,0] = T['g'][0,2] = 2*x*(z-3*y)/(r**2*np.sin(theta))
    T['g'][1,1] = 6*x**2/(r**2*np.sin(theta)**2) - (6*x**2+4*y*z)/r**2
    T['g'][1,0] = T['g'][0,1] = -2*x*(x**2+y**2+3*y*z)/(r**3*np.sin(theta)**2)
    T['g'][0,0] = 6*y**2/(x**2+y**2)
    v = d3.DotProduct(T, u).evaluate()                                   # test_dot_product_tensor_vector
    assert np.allclose(at(v), np.sum(at(T)*at(u)[:,None,:,:,:], axis=0))
    f = d.Field(bases=b); f.preset_scales(dealias); f['g'] = x**3 + 2*y**3 + 3*z**3
    hg = x**3 + 2*y**3 + 3*z**3
    assert np.allclose(at((2 * f).evaluate()), 2*hg)                     # test_multiply_number_scalar
    assert np.allclose(at((f * 2).evaluate()), 2*hg)                     # test_multiply_scalar_number
    assert np.allclose(at((f * f).evaluate()), hg**2)                    # test_multiply_scalar_scalar
    u = d3.Gradient(f, c).evaluate()
    v = (f * u).evaluate()                                               # test_multiply_scalar_vector
    assert np.allclose(at(v), at(f)[None,...]*at(u))
    Tt = (u * u).evaluate()                                              # test_multiply_vector_vector
    assert np.allclose(at(Tt), at(u)[None,...] * at(u)[:,None,...])
    G = d3.Gradient(u, c).evaluate()
    Q = (u * G).evaluate()                                               # test_multiply_vector_tensor
    assert np.allclose(at(Q), at(u)[:,None,None,...] * at(G)[None,...])
    Q = (G * G).evaluate()                                               # test_multiply_tensor_tensor
    assert np.allclose(at(Q), at(G)[:,:,None,None,...] * at(G)[None,None,...])
