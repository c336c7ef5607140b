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
