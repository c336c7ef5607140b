"""The reference's CFL tests (dedalus/tests/test_cfl.py:59-228) restated against `dedalus_b200`, real dtype: the AdvectiveCFL operator on
Fourier, Chebyshev, Fourier x Chebyshev, sphere and shell bases, and the CFL tool end to end."""
import numpy as np
import dedalus_b200 as d3

dtype = np.float64


def check_full_cfl_fourier_chebyshev(dealias, safety, Nx=32, Nz=16):
    """test_cfl.py:59-93"""
    Lx, Lz = 2, 1
    c = d3.CartesianCoordinates('x', 'z')
    d = d3.Distributor(c, dtype=dtype)
    xb = d3.Fourier(c.coords[0], size=Nx, bounds=(0, Lx), dealias=dealias, dtype=dtype)
    zb = d3.Chebyshev(c.coords[1], size=Nz, bounds=(0, Lz), dealias=dealias)
    u = d.VectorField(c, bases=(xb, zb))
    problem = d3.IVP([u], namespace=locals())
    problem.add_equation("dt(u) = 0")
    solver = problem.build_solver(d3.SBDF1)
    cfl = d3.CFL(solver, initial_dt=1, safety=safety, cadence=1)
    cfl.add_velocity(u)
    u.fill_random(layout='g')
    for i in range(2):
        solver.step(1)
    dt_cfl = cfl.compute_timestep()
    cfl_op = d3.AdvectiveCFL(u, c)
    cfl_freq = np.abs(u['g'][0] / cfl_op.cfl_spacing()[0])
    cfl_freq += np.abs(u['g'][1] / cfl_op.cfl_spacing()[1])
    cfl_freq = np.max(cfl_freq)
    dt_target = safety / cfl_freq
    assert np.allclose(dt_cfl, dt_target)


def check_cfl_1d(kind, dealias, N=32, L=1.44):
    """test_cfl.py:96-129"""
    c = d3.CartesianCoordinates('x')
    d = d3.Distributor(c, dtype=dtype)
    if kind == 'fourier':
        b = d3.Fourier(c.coords[0], size=N, bounds=(0, L), dealias=dealias, dtype=dtype)
    else:
        b = d3.Chebyshev(c.coords[0], size=N, bounds=(0, L), dealias=dealias)
    u = d.VectorField(c, bases=b)
    u.fill_random(layout='g')
    cfl = d3.AdvectiveCFL(u, c)
    cfl_freq = cfl.evaluate()['g']
    target_freq = np.abs(u['g']) / cfl.cfl_spacing()[0]
    assert np.allclose(cfl_freq, target_freq)


def check_cfl_fourier_chebyshev(dealias, Nx=32, Nz=16):
    """test_cfl.py:132-149"""
    c = d3.CartesianCoordinates('x', 'z')
    d = d3.Distributor(c, dtype=dtype)
    xb = d3.Fourier(c.coords[0], size=Nx, bounds=(0, 2), dealias=dealias, dtype=dtype)
    zb = d3.Chebyshev(c.coords[1], size=Nz, bounds=(0, 1), dealias=dealias)
    u = d.VectorField(c, bases=(xb, zb))
    u.fill_random(layout='g')
    cfl = d3.AdvectiveCFL(u, c)
    cfl_freq = cfl.evaluate()['g']
    target_freq = np.abs(u['g'][0]) / cfl.cfl_spacing()[0]
    target_freq += np.abs(u['g'][1]) / cfl.cfl_spacing()[1]
    assert np.allclose(cfl_freq, target_freq)


def check_cfl_sphere(dealias, N=16):
    """test_cfl.py:152-163"""
    c = d3.S2Coordinates('phi', 'theta')
    d = d3.Distributor(c, dtype=dtype)
    b = d3.SphereBasis(c, (2*N, N), radius=2.5, dealias=dealias, dtype=dtype)
    u = d.VectorField(c, bases=b)
    u.fill_random(layout='g')
    cfl = d3.AdvectiveCFL(u, c)
    cfl_freq = cfl.evaluate()['g']
    target_freq = np.sqrt(u['g'][0]**2 + u['g'][1]**2) / cfl.cfl_spacing()[0]
    assert np.allclose(cfl_freq, target_freq)


def check_cfl_shell(dealias, N=8):
    """test_cfl.py:213-228"""
    c = d3.SphericalCoordinates('phi', 'theta', 'r')
    d = d3.Distributor(c, dtype=dtype)
    b = d3.ShellBasis(c, (2*N, N, N), radii=(0.4, 2.5), dealias=dealias, dtype=dtype)
    u = d.VectorField(c, bases=b)
    u.fill_random(layout='g')
    cfl = d3.AdvectiveCFL(u, c)
    cfl_freq = cfl.evaluate()['g']
    target_freq = np.sqrt(u['g'][0]**2 + u['g'][1]**2) / cfl.cfl_spacing()[0]
    target_freq += np.abs(u['g'][2]) / cfl.cfl_spacing()[1]
    assert np.allclose(cfl_freq, target_freq)
