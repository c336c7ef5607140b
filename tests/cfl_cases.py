"""CFL checks modelled on the reference's dedalus/tests/test_cfl.py:59-228 (real dtype): the advective CFL frequency field of a random
velocity must equal  sum_i |u_i| / spacing_i  (Cartesian) or  sqrt(u_phi^2 + u_theta^2) / s_angular + |u_r| / s_radial  (sphere,
shell) built from the operator's own spacings, and the CFL tool must return safety / max(frequency)."""
import numpy as np
import dedalus_b200 as d3

REAL = np.float64


def _random_velocity(dist, coordsys, bases):
    vel = dist.VectorField(coordsys, bases=bases)
    vel.fill_random(layout='g')
    return vel


def _frequency_and_expectation(vel, coordsys, curvilinear=False):
    op = d3.AdvectiveCFL(vel, coordsys)
    measured = op.evaluate()['g']
    grid = vel['g']                      # evaluating the operator leaves the velocity on the dealiased grid scales, as in the reference
    spacings = op.cfl_spacing()
    if curvilinear:
        expected = np.hypot(grid[0], grid[1]) / spacings[0]
        if len(spacings) > 1:
            expected = expected + np.abs(grid[2]) / spacings[1]
    else:
        expected = sum(np.abs(grid[i]) / spacings[i] for i in range(len(spacings)))
    return measured, expected


def check_cfl_1d(kind, dealias, N=32, L=1.44):
    """test_cfl.py:96-129 -- one periodic or one Chebyshev direction"""
    cs = d3.CartesianCoordinates('x')
    dist = d3.Distributor(cs, dtype=REAL)
    if kind == 'fourier':
        basis = d3.Fourier(cs.coords[0], size=N, bounds=(0, L), dealias=dealias, dtype=REAL)
    else:
        basis = d3.Chebyshev(cs.coords[0], size=N, bounds=(0, L), dealias=dealias)
    got, want = _frequency_and_expectation(_random_velocity(dist, cs, basis), cs)
    assert np.allclose(got, want)


def check_cfl_fourier_chebyshev(dealias, Nx=32, Nz=16):
    """test_cfl.py:132-149"""
    cs = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(cs, dtype=REAL)
    bases = (d3.Fourier(cs.coords[0], size=Nx, bounds=(0, 2), dealias=dealias, dtype=REAL),
             d3.Chebyshev(cs.coords[1], size=Nz, bounds=(0, 1), dealias=dealias))
    got, want = _frequency_and_expectation(_random_velocity(dist, cs, bases), cs)
    assert np.allclose(got, want)


def check_cfl_sphere(dealias, N=16):
    """test_cfl.py:152-163"""
    cs = d3.S2Coordinates('phi', 'theta')
    dist = d3.Distributor(cs, dtype=REAL)
    basis = d3.SphereBasis(cs, (2 * N, N), radius=2.5, dealias=dealias, dtype=REAL)
    got, want = _frequency_and_expectation(_random_velocity(dist, cs, basis), cs, curvilinear=True)
    assert np.allclose(got, want)


def check_cfl_shell(dealias, N=8):
    """test_cfl.py:213-228"""
    cs = d3.SphericalCoordinates('phi', 'theta', 'r')
    dist = d3.Distributor(cs, dtype=REAL)
    basis = d3.ShellBasis(cs, (2 * N, N, N), radii=(0.4, 2.5), dealias=dealias, dtype=REAL)
    got, want = _frequency_and_expectation(_random_velocity(dist, cs, basis), cs, curvilinear=True)
    assert np.allclose(got, want)


def check_full_cfl_fourier_chebyshev(dealias, safety, Nx=32, Nz=16):
    """test_cfl.py:59-93 -- the tool end to end: a frozen random flow (dt(u) = 0), two unit steps, then the proposed time step"""
    cs = d3.CartesianCoordinates('x', 'z')
    dist = d3.Distributor(cs, dtype=REAL)
    bases = (d3.Fourier(cs.coords[0], size=Nx, bounds=(0, 2), dealias=dealias, dtype=REAL),
             d3.Chebyshev(cs.coords[1], size=Nz, bounds=(0, 1), dealias=dealias))
    flow = dist.VectorField(cs, bases=bases)
    frozen = d3.IVP([flow], namespace=dict(u=flow))
    frozen.add_equation("dt(u) = 0")
    stepper = frozen.build_solver(d3.SBDF1)
    tool = d3.CFL(stepper, initial_dt=1, safety=safety, cadence=1)
    tool.add_velocity(flow)
    flow.fill_random(layout='g')
    stepper.step(1)
    stepper.step(1)
    proposed = tool.compute_timestep()
    spacings = d3.AdvectiveCFL(flow, cs).cfl_spacing()
    values = flow['g']
    peak = np.max(np.abs(values[0] / spacings[0]) + np.abs(values[1] / spacings[1]))
    assert np.allclose(proposed, safety / peak)
