"""Shared bodies of the complex-dtype (T3) tests: run under the CPU emulation (tests/test_emu_complex.py) and on the GPU
(tests/test_gpu_t3_complex.py)."""
import numpy as np
import dedalus_b200 as d3
from dedalus_b200 import examples

TOL = dict(rtol=1e-8, atol=1e-12)


def check_heat_periodic(timestepper):
    """The reference's own complex IVP test, tests/test_ivp.py:18-49: 1-D heat equation on ComplexFourier in complex128, every
    scheme, 20 steps, analytic solution, same assertion (np.allclose defaults)."""
    from dedalus_b200 import timesteppers as ts
    scheme = ts.schemes.get(timestepper, getattr(ts, timestepper, None))
    c = d3.Coordinate('x')
    d = d3.Distributor(c, dtype=np.complex128)
    b = d3.ComplexFourier(c, size=8, bounds=(0, 2 * np.pi), dealias=1)
    x = d.local_grid(b, scale=1)
    u = d.Field(bases=b); F = d.Field(bases=b)
    F['g'] = np.sin(x)
    dx = lambda A: d3.Differentiate(A, c)
    problem = d3.IVP([u], namespace=locals())
    problem.add_equation("dt(u) - dx(dx(u)) = F")
    solver = problem.build_solver(scheme)
    for i in range(20):
        solver.step(1e-5)
    amp = 1 - np.exp(-solver.sim_time)
    u.change_scales(1)
    assert u['g'].dtype == np.complex128
    assert np.allclose(u['g'], amp * np.sin(x))


def check_ginzburg_landau(g, tag, scheme):
    """ComplexFourier x ChebyshevT, complex LHS coefficients, advective + cubic RHS, tau terms: K steps vs the reference state."""
    Nx, Nz, steps, dt = (float(v) for v in g[f"{tag}_meta"])
    pb = examples.complex_ginzburg_landau(int(Nx), int(Nz))
    solver = pb['problem'].build_solver(getattr(d3, scheme))
    examples.complex_ginzburg_landau_initial_condition(pb['u'], pb['bases'])
    assert np.allclose(pb['u']['c'], g[f"{tag}_u0"], rtol=1e-12, atol=1e-15)
    for _ in range(int(steps)):
        solver.step(dt)
    for name, key in (('u', 'u1'), ('tau1', 'tau1'), ('tau2', 'tau2')):
        ref = g[f"{tag}_{key}"]
        assert np.allclose(pb[name]['c'], ref, **TOL), (name, np.abs(pb[name]['c'] - ref).max())
    return solver
