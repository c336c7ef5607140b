"""Problem builders for the BASELINE configs, written against the public API exactly like the stock scripts
(examples/ivp_1d_kdv_burgers/kdv_burgers.py:22-54, examples/ivp_2d_rayleigh_benard/rayleigh_benard.py:33-89,
and the 3-D extension defined in SURVEY.md Appendix C)."""
import numpy as np
import dedalus_b200 as d3


def kdv_burgers(N=1024, Lx=10, a=1e-4, b=2e-4, dealias=3/2, dtype=np.float64):
    xcoord = d3.Coordinate('x')
    dist = d3.Distributor(xcoord, dtype=dtype)
    xbasis = d3.RealFourier(xcoord, size=N, bounds=(0, Lx), dealias=dealias)
    u = dist.Field(name='u', bases=xbasis)
    dx = lambda A: d3.Differentiate(A, xcoord)
    problem = d3.IVP([u], namespace=locals())
    problem.add_equation("dt(u) - a*dx(dx(u)) - b*dx(dx(dx(u))) = - u*dx(u)")
    return dict(problem=problem, dist=dist, u=u, xbasis=xbasis, Lx=Lx)


def kdv_initial_condition(u, xbasis, Lx, n=20):
    x = u.dist.local_grid(xbasis)
    u['g'] = np.log(1 + np.cosh(n)**2 / np.cosh(n * (x - 0.2 * Lx))**2) / (2 * n)


def rayleigh_benard(dim=3, Nh=256, Nz=256, Rayleigh=1e6, Prandtl=1, Lx=4, Lz=1, dealias=3/2, mesh=None, dtype=np.float64, Nx=None):
    names = ('x', 'z') if dim == 2 else ('x', 'y', 'z')
    coords = d3.CartesianCoordinates(*names)
    dist = d3.Distributor(coords, dtype=dtype, mesh=mesh)
    # Nx: size of the FIRST horizontal axis if different from Nh (slab-shaped problems: one rank's share of a distributed run)
    hb = tuple(d3.RealFourier(coords[n], size=(Nx if (Nx and i == 0) else Nh), bounds=(0, Lx), dealias=dealias) for i, n in enumerate(names[:-1]))
    zb = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, Lz), dealias=dealias)
    bases = hb + (zb,)
    p = dist.Field(name='p', bases=bases)
    b = dist.Field(name='b', bases=bases)
    u = dist.VectorField(coords, name='u', bases=bases)
    tau_p = dist.Field(name='tau_p')
    tau_b1 = dist.Field(name='tau_b1', bases=hb)
    tau_b2 = dist.Field(name='tau_b2', bases=hb)
    tau_u1 = dist.VectorField(coords, name='tau_u1', bases=hb)
    tau_u2 = dist.VectorField(coords, name='tau_u2', bases=hb)
    kappa = (Rayleigh * Prandtl)**(-1/2)
    nu = (Rayleigh / Prandtl)**(-1/2)
    ez = coords.unit_vector_fields(dist)[-1]
    lift_basis = zb.derivative_basis(1)
    lift = lambda A: d3.Lift(A, lift_basis, -1)
    grad_u = d3.grad(u) + ez * lift(tau_u1)
    grad_b = d3.grad(b) + ez * lift(tau_b1)
    problem = d3.IVP([p, b, u, tau_p, tau_b1, tau_b2, tau_u1, tau_u2], namespace=locals())
    problem.add_equation("trace(grad_u) + tau_p = 0")
    problem.add_equation("dt(b) - kappa*div(grad_b) + lift(tau_b2) = - u@grad(b)")
    problem.add_equation("dt(u) - nu*div(grad_u) + grad(p) - b*ez + lift(tau_u2) = - u@grad(u)")
    problem.add_equation("b(z=0) = Lz")
    problem.add_equation("u(z=0) = 0")
    problem.add_equation("b(z=Lz) = 0")
    problem.add_equation("u(z=Lz) = 0")
    problem.add_equation("integ(p) = 0")
    return dict(problem=problem, dist=dist, coords=coords, p=p, b=b, u=u, bases=bases, zb=zb, Lz=Lz,
                taus=(tau_p, tau_b1, tau_b2, tau_u1, tau_u2))


def rayleigh_benard_initial_condition(b, bases, Lz, seed=42):
    dist = b.dist
    z = dist.local_grids(*bases)[-1]
    b.fill_random('g', seed=seed, distribution='normal', scale=1e-3)
    b['g'] *= z * (Lz - z)
    b['g'] += Lz - z


def shallow_water(Nphi=512, Ntheta=256, dealias=3/2, dtype=np.float64):
    """Spherical shallow water, BASELINE config 4 (examples/ivp_sphere_shallow_water/shallow_water.py:25-86)."""
    meter = 1 / 6.37122e6
    hour = 1
    second = hour / 3600
    R = 6.37122e6 * meter
    Omega = 7.292e-5 / second
    nu = 1e5 * meter**2 / second / 32**2
    g = 9.80616 * meter / second**2
    H = 1e4 * meter
    timestep = 600 * second * min(1.0, 128 / Ntheta)
    coords = d3.S2Coordinates('phi', 'theta')
    dist = d3.Distributor(coords, dtype=dtype)
    basis = d3.SphereBasis(coords, (Nphi, Ntheta), radius=R, dealias=dealias, dtype=dtype)
    u = dist.VectorField(coords, name='u', bases=basis)
    h = dist.Field(name='h', bases=basis)
    zcross = lambda A: d3.MulCosine(d3.skew(A))
    problem = d3.IVP([u, h], namespace=locals())
    problem.add_equation("dt(u) + nu*lap(lap(u)) + g*grad(h) + 2*Omega*zcross(u) = - u@grad(u)")
    problem.add_equation("dt(h) + nu*lap(lap(h)) + H*div(u) = - div(h*u)")
    return dict(problem=problem, dist=dist, coords=coords, basis=basis, u=u, h=h, timestep=timestep,
                units=dict(meter=meter, hour=hour, second=second))


def shallow_water_initial_condition(u, h, basis, units):
    """Zonal jet and height perturbation of the stock script (lines 48-57, 68-73), without the balancing LBVP."""
    meter, second = units['meter'], units['second']
    phi, theta = u.dist.local_grids(basis)
    lat = np.pi / 2 - theta + 0 * phi
    umax = 80 * meter / second
    lat0 = np.pi / 7
    lat1 = np.pi / 2 - lat0
    en = np.exp(-4 / (lat1 - lat0)**2)
    jet = (lat0 <= lat) * (lat <= lat1)
    u_jet = umax / en * np.exp(1 / (lat[jet] - lat0) / (lat[jet] - lat1))
    u['g'][0][jet] = u_jet
    lat2 = np.pi / 4
    hpert = 120 * meter
    alpha = 1 / 3
    beta = 1 / 15
    h['g'] += hpert * np.cos(lat) * np.exp(-(phi / alpha)**2) * np.exp(-((lat2 - lat) / beta)**2)


def complex_ginzburg_landau(Nx=16, Nz=12, dealias=3/2):
    """Complex-dtype test problem (T3): ComplexFourier x ChebyshevT in complex128 with complex LHS coefficients, an advective and a
    cubic term (tests/golden/make_golden.py complex_cgl runs the same script through the reference)."""
    coords = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(coords, dtype=np.complex128)
    xb = d3.ComplexFourier(coords['x'], size=Nx, bounds=(0, 2 * np.pi), dealias=dealias)
    zb = d3.ChebyshevT(coords['z'], size=Nz, bounds=(0, 1), dealias=dealias)
    u = dist.Field(name='u', bases=(xb, zb))
    tau1 = dist.Field(name='tau1', bases=xb)
    tau2 = dist.Field(name='tau2', bases=xb)
    lift_basis = zb.derivative_basis(2)
    lift = lambda A, n: d3.Lift(A, lift_basis, n)
    dx = lambda A: d3.Differentiate(A, coords['x'])
    c1, c2 = 0.3 + 0.2j, 1.0 - 0.5j
    problem = d3.IVP([u, tau1, tau2], namespace=locals())
    problem.add_equation("dt(u) - c1*lap(u) + (0.5j)*dx(u) + lift(tau1,-1) + lift(tau2,-2) = - c2*u*dx(u) + u*u*u")
    problem.add_equation("u(z=0) = 0")
    problem.add_equation("u(z=1) = 0")
    return dict(problem=problem, dist=dist, u=u, tau1=tau1, tau2=tau2, bases=(xb, zb))


def complex_ginzburg_landau_initial_condition(u, bases):
    x, z = u.dist.local_grids(*bases)
    u.fill_random('g', seed=3, distribution='normal', scale=0.1)
    u['g'] *= z * (1 - z)


def shell_convection(Nphi=256, Ntheta=128, Nr=128, Ri=14, Ro=15, Rayleigh=3500, Prandtl=1, dealias=3/2, dtype=np.float64, rhs_b_extra=None):
    """Boussinesq convection in a spherical shell, BASELINE config 5 (examples/ivp_shell_convection/shell_convection.py:33-83)."""
    coords = d3.SphericalCoordinates('phi', 'theta', 'r')
    dist = d3.Distributor(coords, dtype=dtype)
    shell = d3.ShellBasis(coords, shape=(Nphi, Ntheta, Nr), radii=(Ri, Ro), dealias=dealias, dtype=dtype)
    sphere = shell.outer_surface
    p = dist.Field(name='p', bases=shell)
    b = dist.Field(name='b', bases=shell)
    u = dist.VectorField(coords, name='u', bases=shell)
    tau_p = dist.Field(name='tau_p')
    tau_b1 = dist.Field(name='tau_b1', bases=sphere)
    tau_b2 = dist.Field(name='tau_b2', bases=sphere)
    tau_u1 = dist.VectorField(coords, name='tau_u1', bases=sphere)
    tau_u2 = dist.VectorField(coords, name='tau_u2', bases=sphere)
    kappa = (Rayleigh * Prandtl)**(-1/2)
    nu = (Rayleigh / Prandtl)**(-1/2)
    phi, theta, r = dist.local_grids(shell)
    er = dist.VectorField(coords, bases=shell.radial_basis)
    er['g'][2] = 1
    rvec = dist.VectorField(coords, bases=shell.radial_basis)
    rvec['g'][2] = r
    lift_basis = shell.derivative_basis(1)
    lift = lambda A: d3.Lift(A, lift_basis, -1)
    grad_u = d3.grad(u) + rvec*lift(tau_u1)
    grad_b = d3.grad(b) + rvec*lift(tau_b1)
    problem = d3.IVP([p, b, u, tau_p, tau_b1, tau_b2, tau_u1, tau_u2], namespace=locals())
    problem.add_equation("trace(grad_u) + tau_p = 0")
    problem.add_equation("dt(b) - kappa*div(grad_b) + lift(tau_b2) = - u@grad(b)" + (" + " + rhs_b_extra if rhs_b_extra else ""))
    problem.add_equation("dt(u) - nu*div(grad_u) + grad(p) - b*er + lift(tau_u2) = - u@grad(u)")
    problem.add_equation("b(r=Ri) = 1")
    problem.add_equation("u(r=Ri) = 0")
    problem.add_equation("b(r=Ro) = 0")
    problem.add_equation("u(r=Ro) = 0")
    problem.add_equation("integ(p) = 0")
    return dict(problem=problem, dist=dist, coords=coords, shell=shell, p=p, b=b, u=u, Ri=Ri, Ro=Ro,
                taus=dict(tau_p=tau_p, tau_b1=tau_b1, tau_b2=tau_b2, tau_u1=tau_u1, tau_u2=tau_u2))


def shell_convection_initial_condition(b, shell, Ri, Ro, seed=42):
    """Random noise damped at the walls on top of the conductive profile (stock script lines 86-88)."""
    phi, theta, r = b.dist.local_grids(shell)
    b.fill_random('g', seed=seed, distribution='normal', scale=1e-3)
    b['g'] *= (r - Ri) * (Ro - r)
    b['g'] += (Ri - Ri * Ro / r) / (Ri - Ro)
