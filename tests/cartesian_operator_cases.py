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


def _random(dist, tensorsig, bases, layout=None):
    field = dist.TensorField(tensorsig, bases=bases) if tensorsig else dist.Field(bases=bases)
    field.fill_random(layout='g')
    if layout:
        field.change_layout(layout)
    return field


def _solve(unknowns, lhs, rhs, **names):
    bvp = d3.LBVP(unknowns, namespace=names)
    bvp.add_equation(f"{lhs} = {rhs}")
    bvp.build_solver().solve()


def check_skew(kind):
    """test_cartesian_operators.py:93-132: skew(v) = (-v_y, v_x) in either layout; the LBVP skew(u) = skew(v) returns v."""
    for layout in ('c', 'g'):
        cs, dist, bases, grids = build(kind)
        v = _random(dist, (cs,), bases, layout)
        rotated = d3.skew(v).evaluate()
        assert np.allclose(rotated[layout][0], -v[layout][1]) and np.allclose(rotated[layout][1], v[layout][0])
    cs, dist, bases, grids = build(kind)
    v, u = _random(dist, (cs,), bases), dist.VectorField(cs, bases=bases)
    _solve([u], "skew(u)", "skew(v)", u=u, v=v)
    assert np.allclose(u['c'], v['c'])


def check_trace_and_transpose(kind, N=16):
    """test_cartesian_operators.py:134-250: trace of rank-2 / rank-3 tensors and the transpose against numpy, in both layouts; the
    LBVPs trace(I*u) = dim*f (scalar and vector f) and transpose(u) = transpose(T) return their data."""
    if len(kind) == 3:
        N = 8
    for layout in ('c', 'g'):
        cs, dist, bases, grids = build(kind, N)
        T2 = _random(dist, (cs, cs), bases, layout)
        assert np.allclose(d3.trace(T2).evaluate()[layout], np.trace(T2[layout]))
        T3 = _random(dist, (cs, cs, cs), bases, layout)
        assert np.allclose(d3.trace(T3).evaluate()[layout], np.trace(T3[layout]))
        axes = list(range(2 + len(grids)))
        axes[0], axes[1] = 1, 0
        assert np.allclose(d3.transpose(T2).evaluate()[layout], np.transpose(T2[layout], axes))
    cs, dist, bases, grids = build(kind, N)
    dim = len(grids)
    identity = dist.TensorField((cs, cs))
    for i in range(dim):
        identity['g'][i, i] = 1
    for sig in ((), (cs,)):
        data = _random(dist, sig, bases)
        unknown = dist.TensorField(sig, bases=bases) if sig else dist.Field(bases=bases)
        _solve([unknown], "trace(I*u)", "dim*f", u=unknown, f=data, I=identity, dim=dim)
        assert np.allclose(unknown['c'], data['c'])
    T2 = _random(dist, (cs, cs), bases)
    unknown = dist.TensorField((cs, cs), bases=bases)
    _solve([unknown], "transpose(u)", "transpose(f)", u=unknown, f=T2)
    assert np.allclose(unknown['c'], T2['c'])


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
