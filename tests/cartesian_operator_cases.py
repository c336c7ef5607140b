"""The reference's Cartesian tensor-operator tests (dedalus/tests/test_cartesian_operators.py:93-250) restated against `dedalus_b200`,
real dtype: skew, trace (rank 2 and 3) and transpose, evaluated explicitly in grid and coefficient space and implicitly through LBVPs,
on Fourier x Fourier, Fourier x Chebyshev and their 3-D versions.  (Chebyshev x Chebyshev is outside this build: Jacobi bases are
supported on the last axis.)"""
import numpy as np
import dedalus_b200 as d3

dtype = np.float64
Lx, Ly, Lz = 1.3, 2.4, 1.9


def build(kind, N=16, dealias=1):
    if len(kind) == 2:
        c = d3.CartesianCoordinates('x', 'y')
        d = d3.Distributor(c, dtype=dtype)
        xb = d3.RealFourier(c.coords[0], size=N, bounds=(0, Lx), dealias=dealias)
        yb = (d3.RealFourier if kind[1] == 'F' else d3.Chebyshev)(c.coords[1], size=N, bounds=(0, Ly), dealias=dealias)
        b = (xb, yb)
    else:
        c = d3.CartesianCoordinates('x', 'y', 'z')
        d = d3.Distributor(c, dtype=dtype)
        xb = d3.RealFourier(c.coords[0], size=N, bounds=(0, Lx), dealias=dealias)
        yb = d3.RealFourier(c.coords[1], size=N, bounds=(0, Ly), dealias=dealias)
        zb = (d3.RealFourier if kind[2] == 'F' else d3.ChebyshevT)(c.coords[2], size=N, bounds=(0, Lz), dealias=dealias)
        b = (xb, yb, zb)
    r = d.local_grids(*b, scales=dealias)
    return c, d, b, r


def check_skew(kind):
    for layout in ('c', 'g'):                                            # test_skew_explicit
        c, d, b, r = build(kind)
        f = d.VectorField(c, bases=b)
        f.fill_random(layout='g')
        f.change_layout(layout)
        g = d3.skew(f).evaluate()
        assert np.allclose(g[layout][0], -f[layout][1])
        assert np.allclose(g[layout][1], f[layout][0])
    c, d, b, r = build(kind)                                             # test_skew_implicit
    f = d.VectorField(c, bases=b)
    f.fill_random(layout='g')
    u = d.VectorField(c, bases=b)
    problem = d3.LBVP([u], namespace=locals())
    problem.add_equation("skew(u) = skew(f)")
    solver = problem.build_solver()
    solver.solve()
    assert np.allclose(u['c'], f['c'])


def check_trace_and_transpose(kind, N=16):
    if len(kind) == 3:
        N = 8
    for layout in ('c', 'g'):
        c, d, b, r = build(kind, N)                                      # test_trace_explicit
        f = d.TensorField((c, c), bases=b)
        f.fill_random(layout='g')
        f.change_layout(layout)
        g = d3.trace(f).evaluate()
        assert np.allclose(g[layout], np.trace(f[layout]))
        f3 = d.TensorField((c, c, c), bases=b)                           # test_trace_rank3_explicit
        f3.fill_random(layout='g')
        f3.change_layout(layout)
        g = d3.trace(f3).evaluate()
        assert np.allclose(g[layout], np.trace(f3[layout]))
        g = d3.transpose(f).evaluate()                                   # test_transpose_explicit
        order = np.arange(2 + len(r))
        order[:2] = [1, 0]
        assert np.allclose(g[layout], np.transpose(f[layout], order))
    c, d, b, r = build(kind, N)                                          # test_trace_implicit / test_trace_rank3_implicit
    dim = len(r)
    I = d.TensorField((c, c))
    for i in range(dim):
        I['g'][i, i] = 1
    for make in (lambda: d.Field(bases=b), lambda: d.VectorField(c, bases=b)):
        f = make()
        f.fill_random(layout='g')
        u = make()
        problem = d3.LBVP([u], namespace=dict(u=u, f=f, I=I, dim=dim))
        problem.add_equation("trace(I*u) = dim*f")
        solver = problem.build_solver()
        solver.solve()
        assert np.allclose(u['c'], f['c'])
    f = d.TensorField((c, c), bases=b)                                   # test_transpose_implicit
    f.fill_random(layout='g')
    u = d.TensorField((c, c), bases=b)
    problem = d3.LBVP([u], namespace=locals())
    problem.add_equation("transpose(u) = transpose(f)")
    solver = problem.build_solver()
    solver.solve()
    assert np.allclose(u['c'], f['c'])


def check_curls():
    """test_2d_curl_explicit_vector / _scalar (FF, FC), test_curl_explicit (FFF, FFC), test_curl_implicit_FFC
    (test_cartesian_operators.py:252-400)"""
    N, dealias = 16, 1
    kx, ky = 2*np.pi/Lx, 2*np.pi/Ly
    for kind in ("FF", "FC"):
        c, d, b, (x, y) = build(kind, N, dealias)
        f = d.VectorField(c, bases=b)
        f.preset_scales(dealias)
        f['g'][0] = (np.sin(2*kx*x)+np.sin(kx*x))*np.cos(ky*y)
        f['g'][1] = np.sin(kx*x)*np.cos(ky*y)
        g_op = - d3.div(d3.skew(f))                # z @ curl(f)
        g = d.Field(bases=b)
        g.preset_scales(dealias)
        g['g'] = kx*np.cos(kx*x)*np.cos(ky*y) + ky*(np.sin(2*kx*x)+np.sin(kx*x))*np.sin(ky*y)
        assert np.allclose(g_op.evaluate()['g'], g['g'])
        c, d, b, (x, y) = build(kind, 2*N, dealias)
        f = d.Field(bases=b)
        f.preset_scales(dealias)
        f['g'] = (np.sin(2*kx*x)+np.sin(kx*x))*np.cos(ky*y)
        g_op = - d3.skew(d3.grad(f))               # curl(f*ez)
        g = d.VectorField(c, bases=b)
        g.preset_scales(dealias)
        g['g'][0] = -ky*(np.sin(2*kx*x)+np.sin(kx*x))*np.sin(ky*y)
        g['g'][1] = -(2*kx*np.cos(2*kx*x)+kx*np.cos(kx*x))*np.cos(ky*y)
        assert np.allclose(g_op.evaluate()['g'], g['g'])
    k = 2*np.pi*np.array([1/Lx, 1/Ly, 1/Lz])

    def abc(kind):
        c, d, b, r = build(kind, N, dealias)
        f = d.VectorField(c, bases=b)
        f.preset_scales(dealias)
        f['g'][0] = np.sin(k[2]*r[2]) + np.cos(k[1]*r[1])
        f['g'][1] = np.sin(k[0]*r[0]) + np.cos(k[2]*r[2])
        f['g'][2] = np.sin(k[1]*r[1]) + np.cos(k[0]*r[0])
        g = d.VectorField(c, bases=b)
        g.preset_scales(dealias)
        g['g'][0] = k[2]*np.sin(k[2]*r[2]) + k[1]*np.cos(k[1]*r[1])
        g['g'][1] = k[0]*np.sin(k[0]*r[0]) + k[2]*np.cos(k[2]*r[2])
        g['g'][2] = k[1]*np.sin(k[1]*r[1]) + k[0]*np.cos(k[0]*r[0])
        return c, d, b, r, f, g
    for kind in ("FFF", "FFC"):
        c, d, b, r, f, g = abc(kind)
        assert np.allclose(d3.Curl(f).evaluate()['g'], g['g'])
    c, d, b, r, f, g = abc("FFC")                  # test_curl_implicit_FFC: Helmholtz LBVP
    u = d.VectorField(c, name='u', bases=b)
    phi = d.Field(name='phi', bases=b)
    tau1 = d.VectorField(c, name='tau1', bases=b[0:2])
    tau2 = d.Field(name='tau2', bases=b[0:2])
    lift_basis = b[2].derivative_basis(1)
    lift = lambda A, n: d3.Lift(A, lift_basis, n)
    problem = d3.LBVP([u, phi, tau1, tau2], namespace=locals())
    problem.add_equation("curl(u) + grad(phi) + lift(tau1,-1) = g")
    problem.add_equation("div(u) + lift(tau2,-1) = 0")
    problem.add_equation("u(z=0) = f(z=0)")
    problem.add_equation("phi(z=0) = 0")
    solver = problem.build_solver()
    solver.solve()
    assert np.allclose(u['c'], f['c'])
