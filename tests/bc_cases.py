"""Boundary conditions with data -- lower-dimensional equations whose right-hand side is a field (reference: such an F is carried
in coefficient space like any other, core/problems.py:84-100).  The problem definition is shared by the fixture generator
(tests/golden/make_golden.py gen_bc_data, run with the unmodified reference) and the tests (run with dedalus_b200)."""
import numpy as np


def rb2d_bc_data(d3mod, Nx=16, Nz=16, steps=5, tstep=0.02):
    """2-D Rayleigh-Benard as in examples/ivp_2d_rayleigh_benard/rayleigh_benard.py:44-82 with a non-uniform bottom temperature
    b(z=0) = g(x) and a moving lid u(z=Lz) = (U(x), 0); shared by the generator (reference) and the test (dedalus_b200)."""
    d3 = d3mod
    Lx, Lz = 4, 1
    Rayleigh, Prandtl = 2e5, 1
    coords = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(coords, dtype=np.float64)
    xbasis = d3.RealFourier(coords['x'], size=Nx, bounds=(0, Lx), dealias=3/2)
    zbasis = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, Lz), dealias=3/2)
    p = dist.Field(name='p', bases=(xbasis, zbasis)); b = dist.Field(name='b', bases=(xbasis, zbasis))
    u = dist.VectorField(coords, name='u', bases=(xbasis, zbasis))
    tau_p = dist.Field(name='tau_p')
    tau_b1 = dist.Field(name='tau_b1', bases=xbasis); tau_b2 = dist.Field(name='tau_b2', bases=xbasis)
    tau_u1 = dist.VectorField(coords, name='tau_u1', bases=xbasis); tau_u2 = dist.VectorField(coords, name='tau_u2', bases=xbasis)
    kappa = (Rayleigh * Prandtl)**(-1/2); nu = (Rayleigh / Prandtl)**(-1/2)
    x, z = dist.local_grids(xbasis, zbasis)
    ex, ez = coords.unit_vector_fields(dist)
    lift_basis = zbasis.derivative_basis(1)
    lift = lambda A: d3.Lift(A, lift_basis, -1)
    grad_u = d3.grad(u) + ez*lift(tau_u1)
    grad_b = d3.grad(b) + ez*lift(tau_b1)
    g = dist.Field(name='g', bases=xbasis)
    g['g'] = Lz + 0.2 * np.sin(2 * np.pi * x / Lx) + 0.1 * np.cos(6 * np.pi * x / Lx)
    lid = dist.VectorField(coords, name='lid', bases=xbasis)
    lid['g'][0] = 0.05 * np.cos(2 * np.pi * x / Lx)
    problem = d3.IVP([p, b, u, tau_p, tau_b1, tau_b2, tau_u1, tau_u2], namespace=locals())
    problem.add_equation("trace(grad_u) + tau_p = 0")
    problem.add_equation("dt(b) - kappa*div(grad_b) + lift(tau_b2) = - u@grad(b)")
    problem.add_equation("dt(u) - nu*div(grad_u) + grad(p) - b*ez + lift(tau_u2) = - u@grad(u)")
    problem.add_equation("b(z=0) = g")
    problem.add_equation("u(z=0) = 0")
    problem.add_equation("b(z=Lz) = 0")
    problem.add_equation("u(z=Lz) = 2*lid")
    problem.add_equation("integ(p) = 0")
    solver = problem.build_solver(d3.RK222)
    b.fill_random('g', seed=42, distribution='normal', scale=1e-3)
    b['g'] *= z * (Lz - z)
    b['g'] += Lz - z
    for _ in range(steps):
        solver.step(tstep)
    return dict(p=p, b=b, u=u)




def rb2d_background(d3mod, Nx=16, Nz=16, steps=5, tstep=0.02):
    """Perturbation form of 2-D Rayleigh-Benard: the buoyancy is b0(z) + b with a prescribed background gradient dzB(z) entering
    the right-hand side as a factor WITHOUT an x basis, plus a forcing F(x) without a z basis (broadcast on the grid)."""
    d3 = d3mod
    Lx, Lz = 4, 1
    Rayleigh, Prandtl = 2e5, 1
    coords = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(coords, dtype=np.float64)
    xbasis = d3.RealFourier(coords['x'], size=Nx, bounds=(0, Lx), dealias=3/2)
    zbasis = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, Lz), dealias=3/2)
    p = dist.Field(name='p', bases=(xbasis, zbasis)); b = dist.Field(name='b', bases=(xbasis, zbasis))
    u = dist.VectorField(coords, name='u', bases=(xbasis, zbasis))
    tau_p = dist.Field(name='tau_p')
    tau_b1 = dist.Field(name='tau_b1', bases=xbasis); tau_b2 = dist.Field(name='tau_b2', bases=xbasis)
    tau_u1 = dist.VectorField(coords, name='tau_u1', bases=xbasis); tau_u2 = dist.VectorField(coords, name='tau_u2', bases=xbasis)
    kappa = (Rayleigh * Prandtl)**(-1/2); nu = (Rayleigh / Prandtl)**(-1/2)
    x, z = dist.local_grids(xbasis, zbasis)
    ex, ez = coords.unit_vector_fields(dist)
    lift_basis = zbasis.derivative_basis(1)
    lift = lambda A: d3.Lift(A, lift_basis, -1)
    grad_u = d3.grad(u) + ez*lift(tau_u1)
    grad_b = d3.grad(b) + ez*lift(tau_b1)
    dzB = dist.Field(name='dzB', bases=zbasis)
    dzB['g'] = -1 + 0.3 * z - 0.5 * z**2
    F = dist.Field(name='F', bases=xbasis)
    F['g'] = 0.02 * np.sin(2 * np.pi * x / Lx)
    problem = d3.IVP([p, b, u, tau_p, tau_b1, tau_b2, tau_u1, tau_u2], namespace=locals())
    problem.add_equation("trace(grad_u) + tau_p = 0")
    problem.add_equation("dt(b) - kappa*div(grad_b) + lift(tau_b2) = - u@grad(b) - (u@ez)*dzB + F*b")
    problem.add_equation("dt(u) - nu*div(grad_u) + grad(p) - b*ez + lift(tau_u2) = - u@grad(u)")
    problem.add_equation("b(z=0) = 0")
    problem.add_equation("u(z=0) = 0")
    problem.add_equation("b(z=Lz) = 0")
    problem.add_equation("u(z=Lz) = 0")
    problem.add_equation("integ(p) = 0")
    solver = problem.build_solver(d3.RK222)
    b.fill_random('g', seed=42, distribution='normal', scale=1e-2)
    b['g'] *= z * (Lz - z)
    for _ in range(steps):
        solver.step(tstep)
    return dict(p=p, b=b, u=u)


def rb2d_ncc(d3mod, Nx=16, Nz=16, steps=5, tstep=0.02):
    """Non-constant coefficients on the LEFT-hand side (implicit): a background gradient dzB(z) multiplying the vertical velocity
    and a height-dependent buoyancy factor G(z) -- coefficients varying along the coupled Chebyshev axis (reference NCC matrices,
    core/basis.py:560-628)."""
    d3 = d3mod
    Lx, Lz = 4, 1
    Rayleigh, Prandtl = 2e5, 1
    coords = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(coords, dtype=np.float64)
    xbasis = d3.RealFourier(coords['x'], size=Nx, bounds=(0, Lx), dealias=3/2)
    zbasis = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, Lz), dealias=3/2)
    p = dist.Field(name='p', bases=(xbasis, zbasis)); b = dist.Field(name='b', bases=(xbasis, zbasis))
    u = dist.VectorField(coords, name='u', bases=(xbasis, zbasis))
    tau_p = dist.Field(name='tau_p')
    tau_b1 = dist.Field(name='tau_b1', bases=xbasis); tau_b2 = dist.Field(name='tau_b2', bases=xbasis)
    tau_u1 = dist.VectorField(coords, name='tau_u1', bases=xbasis); tau_u2 = dist.VectorField(coords, name='tau_u2', bases=xbasis)
    kappa = (Rayleigh * Prandtl)**(-1/2); nu = (Rayleigh / Prandtl)**(-1/2)
    x, z = dist.local_grids(xbasis, zbasis)
    ex, ez = coords.unit_vector_fields(dist)
    lift_basis = zbasis.derivative_basis(1)
    lift = lambda A: d3.Lift(A, lift_basis, -1)
    grad_u = d3.grad(u) + ez*lift(tau_u1)
    grad_b = d3.grad(b) + ez*lift(tau_b1)
    dzB = dist.Field(name='dzB', bases=zbasis)
    dzB['g'] = -1 + 0.3 * z - 0.5 * z**2
    G = dist.Field(name='G', bases=zbasis)
    G['g'] = 1 + 0.5 * z
    problem = d3.IVP([p, b, u, tau_p, tau_b1, tau_b2, tau_u1, tau_u2], namespace=locals())
    problem.add_equation("trace(grad_u) + tau_p = 0")
    problem.add_equation("dt(b) - kappa*div(grad_b) + (u@ez)*dzB + lift(tau_b2) = - u@grad(b)")
    problem.add_equation("dt(u) - nu*div(grad_u) + grad(p) - G*b*ez + lift(tau_u2) = - u@grad(u)")
    problem.add_equation("b(z=0) = 0")
    problem.add_equation("u(z=0) = 0")
    problem.add_equation("b(z=Lz) = 0")
    problem.add_equation("u(z=Lz) = 0")
    problem.add_equation("integ(p) = 0")
    solver = problem.build_solver(d3.RK222)
    b.fill_random('g', seed=42, distribution='normal', scale=1e-2)
    b['g'] *= z * (Lz - z)
    for _ in range(steps):
        solver.step(tstep)
    return dict(p=p, b=b, u=u)


def check_ncc(d3, g):
    res = rb2d_ncc(d3)
    for name, f in res.items():
        ref = g["ncc_" + name]
        got = f['c']
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-11 * np.abs(g["ncc_b"]).max()), (name, np.abs(got - ref).max(), np.abs(ref).max())


def rb2d_conservative(d3mod, Nx=16, Nz=16, steps=5, tstep=0.02, mesh=None):
    """Right-hand sides that are not polynomials of derivatives of fields: the advection term in conservative form, div(u*b) -- a
    differential operator applied to a product -- and a grid function, sin(3*b)."""
    d3 = d3mod
    Lx, Lz = 4, 1
    Rayleigh, Prandtl = 2e5, 1
    coords = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(coords, dtype=np.float64, mesh=mesh)
    xbasis = d3.RealFourier(coords['x'], size=Nx, bounds=(0, Lx), dealias=3/2)
    zbasis = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, Lz), dealias=3/2)
    p = dist.Field(name='p', bases=(xbasis, zbasis)); b = dist.Field(name='b', bases=(xbasis, zbasis))
    u = dist.VectorField(coords, name='u', bases=(xbasis, zbasis))
    tau_p = dist.Field(name='tau_p')
    tau_b1 = dist.Field(name='tau_b1', bases=xbasis); tau_b2 = dist.Field(name='tau_b2', bases=xbasis)
    tau_u1 = dist.VectorField(coords, name='tau_u1', bases=xbasis); tau_u2 = dist.VectorField(coords, name='tau_u2', bases=xbasis)
    kappa = (Rayleigh * Prandtl)**(-1/2); nu = (Rayleigh / Prandtl)**(-1/2)
    x, z = dist.local_grids(xbasis, zbasis)
    ex, ez = coords.unit_vector_fields(dist)
    lift_basis = zbasis.derivative_basis(1)
    lift = lambda A: d3.Lift(A, lift_basis, -1)
    grad_u = d3.grad(u) + ez*lift(tau_u1)
    grad_b = d3.grad(b) + ez*lift(tau_b1)
    problem = d3.IVP([p, b, u, tau_p, tau_b1, tau_b2, tau_u1, tau_u2], namespace=locals())
    problem.add_equation("trace(grad_u) + tau_p = 0")
    problem.add_equation("dt(b) - kappa*div(grad_b) + lift(tau_b2) = - div(u*b) + 0.05*sin(3*b)")
    problem.add_equation("dt(u) - nu*div(grad_u) + grad(p) - b*ez + lift(tau_u2) = - u@grad(u)")
    problem.add_equation("b(z=0) = Lz")
    problem.add_equation("u(z=0) = 0")
    problem.add_equation("b(z=Lz) = 0")
    problem.add_equation("u(z=Lz) = 0")
    problem.add_equation("integ(p) = 0")
    solver = problem.build_solver(d3.RK222)
    b.fill_random('g', seed=42, distribution='normal', scale=1e-2)
    b['g'] *= z * (Lz - z)
    b['g'] += Lz - z
    for _ in range(steps):
        solver.step(tstep)
    return dict(p=p, b=b, u=u)


def check_conservative(d3, g):
    res = rb2d_conservative(d3)
    for name, f in res.items():
        ref = g["cons_" + name]
        got = f['c']
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-11 * np.abs(ref).max()), (name, np.abs(got - ref).max(), np.abs(ref).max())


def rb2d_strong(d3mod, Nx=16, Nz=16, steps=5, tstep=0.01):
    """2-D Rayleigh-Benard started with an O(1) velocity field: the advection terms are as large as the linear ones, so the order of
    truncations / conversions in the right-hand side is visible (the noise-started fixtures keep |u| << |b|)."""
    d3 = d3mod
    Lx, Lz = 4, 1
    Rayleigh, Prandtl = 2e5, 1
    coords = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(coords, dtype=np.float64)
    xbasis = d3.RealFourier(coords['x'], size=Nx, bounds=(0, Lx), dealias=3/2)
    zbasis = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, Lz), dealias=3/2)
    p = dist.Field(name='p', bases=(xbasis, zbasis)); b = dist.Field(name='b', bases=(xbasis, zbasis))
    u = dist.VectorField(coords, name='u', bases=(xbasis, zbasis))
    tau_p = dist.Field(name='tau_p')
    tau_b1 = dist.Field(name='tau_b1', bases=xbasis); tau_b2 = dist.Field(name='tau_b2', bases=xbasis)
    tau_u1 = dist.VectorField(coords, name='tau_u1', bases=xbasis); tau_u2 = dist.VectorField(coords, name='tau_u2', bases=xbasis)
    kappa = (Rayleigh * Prandtl)**(-1/2); nu = (Rayleigh / Prandtl)**(-1/2)
    x, z = dist.local_grids(xbasis, zbasis)
    ex, ez = coords.unit_vector_fields(dist)
    lift_basis = zbasis.derivative_basis(1)
    lift = lambda A: d3.Lift(A, lift_basis, -1)
    grad_u = d3.grad(u) + ez*lift(tau_u1)
    grad_b = d3.grad(b) + ez*lift(tau_b1)
    problem = d3.IVP([p, b, u, tau_p, tau_b1, tau_b2, tau_u1, tau_u2], namespace=locals())
    problem.add_equation("trace(grad_u) + tau_p = 0")
    problem.add_equation("dt(b) - kappa*div(grad_b) + lift(tau_b2) = - u@grad(b)")
    problem.add_equation("dt(u) - nu*div(grad_u) + grad(p) - b*ez + lift(tau_u2) = - u@grad(u)")
    problem.add_equation("b(z=0) = Lz")
    problem.add_equation("u(z=0) = 0")
    problem.add_equation("b(z=Lz) = 0")
    problem.add_equation("u(z=Lz) = 0")
    problem.add_equation("integ(p) = 0")
    solver = problem.build_solver(d3.RK222)
    b.fill_random('g', seed=42, distribution='normal', scale=0.3)
    b.low_pass_filter(scales=0.5)
    b['g'] *= z * (Lz - z)
    b['g'] += Lz - z
    u.fill_random('g', seed=43, distribution='normal', scale=0.8)
    u.low_pass_filter(scales=0.5)
    u['g'] *= z * (Lz - z)
    init = dict(b0=np.array(b['c']), u0=np.array(u['c']))
    for _ in range(steps):
        solver.step(tstep)
    return dict(p=p, b=b, u=u), init


def rb3d_strong(d3mod, N=8, steps=3, tstep=0.01, mesh=None):
    """The 3-D problem of the benchmark (Fourier x Fourier x Chebyshev Rayleigh-Benard) started with an O(1)
    velocity field."""
    d3 = d3mod
    Lx, Ly, Lz = 4, 4, 1
    Rayleigh, Prandtl = 1e6, 1
    coords = d3.CartesianCoordinates('x', 'y', 'z')
    dist = d3.Distributor(coords, dtype=np.float64, mesh=mesh)
    xbasis = d3.RealFourier(coords['x'], size=N, bounds=(0, Lx), dealias=3/2)
    ybasis = d3.RealFourier(coords['y'], size=N, bounds=(0, Ly), dealias=3/2)
    zbasis = d3.ChebyshevT(coords['z'], size=N, bounds=(0, Lz), dealias=3/2)
    bases = (xbasis, ybasis, zbasis)
    p = dist.Field(name='p', bases=bases); b = dist.Field(name='b', bases=bases)
    u = dist.VectorField(coords, name='u', bases=bases)
    tau_p = dist.Field(name='tau_p')
    tau_b1 = dist.Field(name='tau_b1', bases=(xbasis, ybasis)); tau_b2 = dist.Field(name='tau_b2', bases=(xbasis, ybasis))
    tau_u1 = dist.VectorField(coords, name='tau_u1', bases=(xbasis, ybasis)); tau_u2 = dist.VectorField(coords, name='tau_u2', bases=(xbasis, ybasis))
    kappa = (Rayleigh * Prandtl)**(-1/2); nu = (Rayleigh / Prandtl)**(-1/2)
    x, y, z = dist.local_grids(xbasis, ybasis, zbasis)
    ex, ey, ez = coords.unit_vector_fields(dist)
    lift_basis = zbasis.derivative_basis(1)
    lift = lambda A: d3.Lift(A, lift_basis, -1)
    grad_u = d3.grad(u) + ez*lift(tau_u1)
    grad_b = d3.grad(b) + ez*lift(tau_b1)
    problem = d3.IVP([p, b, u, tau_p, tau_b1, tau_b2, tau_u1, tau_u2], namespace=locals())
    problem.add_equation("trace(grad_u) + tau_p = 0")
    problem.add_equation("dt(b) - kappa*div(grad_b) + lift(tau_b2) = - u@grad(b)")
    problem.add_equation("dt(u) - nu*div(grad_u) + grad(p) - b*ez + lift(tau_u2) = - u@grad(u)")
    problem.add_equation("b(z=0) = Lz")
    problem.add_equation("u(z=0) = 0")
    problem.add_equation("b(z=Lz) = 0")
    problem.add_equation("u(z=Lz) = 0")
    problem.add_equation("integ(p) = 0")
    solver = problem.build_solver(d3.RK222)
    b.fill_random('g', seed=42, distribution='normal', scale=0.3)
    b.low_pass_filter(scales=0.5)
    b['g'] *= z * (Lz - z)
    b['g'] += Lz - z
    u.fill_random('g', seed=43, distribution='normal', scale=0.8)
    u.low_pass_filter(scales=0.5)
    u['g'] *= z * (Lz - z)
    init = dict(b0=np.array(b['c']), u0=np.array(u['c']))
    for _ in range(steps):
        solver.step(tstep)
    return dict(p=p, b=b, u=u), init


def check_strong_3d(d3, g):
    res, init = rb3d_strong(d3)
    for name, f in res.items():
        ref = g["strong3d_" + name]
        got = f['c']
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-11 * np.abs(ref).max()), (name, np.abs(got - ref).max(), np.abs(ref).max())


def check_strong(d3, g):
    res, init = rb2d_strong(d3)
    for name, f in res.items():
        ref = g["strong_" + name]
        got = f['c']
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-11 * np.abs(ref).max()), (name, np.abs(got - ref).max(), np.abs(ref).max())


def check_background(d3, g):
    res = rb2d_background(d3)
    for name, f in res.items():
        ref = g["bg_" + name]
        got = f['c']
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-11 * np.abs(g["bg_b"]).max()), (name, np.abs(got - ref).max(), np.abs(ref).max())


def check_bc_data(d3, g):
    res = rb2d_bc_data(d3)
    for name, f in res.items():
        ref = g[name]
        got = f['c']
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-11 * np.abs(ref).max()), (name, np.abs(got - ref).max(), np.abs(ref).max())


def check_poisson_lbvp(d3, g):
    """examples/lbvp_2d_poisson/poisson.py:25-65 at 32 x 16: randomly forced Poisson equation with mixed boundary data, against the
    reference executing the stock script (tests/golden/stock_scripts.npz, tag poisson)."""
    Lx, Ly = 2*np.pi, np.pi
    Nx, Ny = 32, 16
    coords = d3.CartesianCoordinates('x', 'y')
    dist = d3.Distributor(coords, dtype=np.float64)
    xbasis = d3.RealFourier(coords['x'], size=Nx, bounds=(0, Lx))
    ybasis = d3.Chebyshev(coords['y'], size=Ny, bounds=(0, Ly))
    u = dist.Field(name='u', bases=(xbasis, ybasis))
    tau_1 = dist.Field(name='tau_1', bases=xbasis)
    tau_2 = dist.Field(name='tau_2', bases=xbasis)
    x, y = dist.local_grids(xbasis, ybasis)
    f = dist.Field(bases=(xbasis, ybasis))
    g_ = dist.Field(bases=xbasis)
    h = dist.Field(bases=xbasis)
    f.fill_random('g', seed=40)
    f.low_pass_filter(shape=(16, 8))
    g_['g'] = np.sin(8*x) * 0.025
    h['g'] = 0
    dy = lambda A: d3.Differentiate(A, coords['y'])
    lift_basis = ybasis.derivative_basis(2)
    lift = lambda A, n: d3.Lift(A, lift_basis, n)
    problem = d3.LBVP([u, tau_1, tau_2], namespace=dict(u=u, tau_1=tau_1, tau_2=tau_2, f=f, g=g_, h=h, dy=dy, lift=lift, Ly=Ly))
    problem.add_equation("lap(u) + lift(tau_1,-1) + lift(tau_2,-2) = f")
    problem.add_equation("u(y=0) = g")
    problem.add_equation("dy(u)(y=Ly) = h")
    solver = problem.build_solver()
    solver.solve()
    for name, fld in (("u", u), ("tau_1", tau_1), ("tau_2", tau_2)):
        ref = g[f"poisson_{name}"]
        assert np.allclose(fld['c'], ref, rtol=1e-9, atol=1e-12 * np.abs(g["poisson_u"]).max()), (name, np.abs(fld['c'] - ref).max())


def shallow_water_forced(d3mod, Nphi=32, Ntheta=16, steps=3):
    """Shallow water on the sphere (examples/ivp_sphere_shallow_water/shallow_water.py:36-86, analytic jet) with a grid function and
    a prescribed forcing field on the right-hand side of the height equation: outside what the fused sphere plan covers, so the
    general evaluator carries the right-hand sides."""
    d3 = d3mod
    meter = 1 / 6.37122e6; hour = 1; second = hour / 3600
    R = 6.37122e6 * meter; Omega = 7.292e-5 / second; nu = 1e5 * meter**2 / second / 32**2
    g = 9.80616 * meter / second**2; H = 1e4 * meter; timestep = 600 * second * min(1.0, 128 / Ntheta)
    coords = d3.S2Coordinates('phi', 'theta')
    dist = d3.Distributor(coords, dtype=np.float64)
    basis = d3.SphereBasis(coords, (Nphi, Ntheta), radius=R, dealias=3/2, dtype=np.float64)
    u = dist.VectorField(coords, name='u', bases=basis)
    h = dist.Field(name='h', bases=basis)
    Q = dist.Field(name='Q', bases=basis)
    zcross = lambda A: d3.MulCosine(d3.skew(A))
    phi, theta = dist.local_grids(basis)
    lat = np.pi / 2 - theta + 0*phi
    umax = 80 * meter / second
    lat0 = np.pi / 7; lat1 = np.pi / 2 - lat0
    en = np.exp(-4 / (lat1 - lat0)**2)
    jet = (lat0 <= lat) * (lat <= lat1)
    u_jet = umax / en * np.exp(1 / (lat[jet] - lat0) / (lat[jet] - lat1))
    u['g'][0][jet] = u_jet
    lat2 = np.pi / 4; hpert = 120 * meter; alpha = 1 / 3; beta = 1 / 15
    h['g'] += hpert * np.cos(lat) * np.exp(-(phi/alpha)**2) * np.exp(-((lat2-lat)/beta)**2)
    Q['g'] = 1e-6 * np.cos(lat)**2 * np.sin(2*phi)
    amp, scale = 2e-6, 3e4
    problem = d3.IVP([u, h], namespace=locals())
    problem.add_equation("dt(u) + nu*lap(lap(u)) + g*grad(h) + 2*Omega*zcross(u) = - u@grad(u)")
    problem.add_equation("dt(h) + nu*lap(lap(h)) + H*div(u) = - div(h*u) + amp*sin(scale*h) + Q")
    solver = problem.build_solver(d3.RK222)
    for _ in range(steps):
        solver.step(timestep)
    return dict(u=u, h=h), solver


def check_shallow_water_forced(d3, g):
    res, solver = shallow_water_forced(d3)
    assert type(solver.rhs_plan).__name__ == "GenericCurvilinearRHS"
    for name, f in res.items():
        ref = g["swf_" + name]
        got = f['c']
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-12 * np.abs(ref).max()), (name, np.abs(got - ref).max(), np.abs(ref).max())
