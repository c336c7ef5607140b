"""The Cartesian cases of the reference's own LBVP tests (dedalus/tests/test_lbvp.py:38-111), real dtype, restated against
`dedalus_b200`: the expected answers are analytic, exactly as in the reference's tests (np.allclose with numpy's defaults)."""
import numpy as np
import dedalus_b200 as d3

dtype = np.float64


def check_algebraic():
    """test_lbvp.py:38-55"""
    coord = d3.Coordinate('x')
    dist = d3.Distributor(coord, dtype=dtype)
    u = dist.Field(name='u')
    v = dist.Field(name='v')
    F = dist.Field(name='F')
    v['g'] = -1
    F['g'] = -3
    problem = d3.LBVP([u], namespace=locals())
    problem.add_equation("v*u = F")
    solver = problem.build_solver()
    solver.solve()
    assert np.allclose(u['g'], 3)


def check_poisson_fourier(N=32):
    """test_lbvp.py:58-82"""
    coord = d3.Coordinate('x')
    dist = d3.Distributor(coord, dtype=dtype)
    basis = d3.Fourier(coord, size=N, bounds=(0, 2*np.pi), dtype=dtype)
    x = dist.local_grid(basis)
    u = dist.Field(name='u', bases=basis)
    g = dist.Field(name='c')
    u_true = np.sin(x)
    f = dist.Field(bases=basis)
    f['g'] = -np.sin(x)
    dx = lambda A: d3.Differentiate(A, coord)
    integ = lambda A: d3.Integrate(A, coord)
    problem = d3.LBVP([u, g], namespace=locals())
    problem.add_equation("dx(dx(u)) + g = f")
    problem.add_equation("integ(u) = 0")
    solver = problem.build_solver()
    solver.solve()
    assert np.allclose(u['g'], u_true)


def check_poisson_jacobi(a, b, N=32):
    """test_lbvp.py:85-111"""
    coord = d3.Coordinate('x')
    dist = d3.Distributor(coord, dtype=dtype)
    basis = d3.Jacobi(coord, size=N, bounds=(0, 2*np.pi), a=a, b=b)
    x = dist.local_grid(basis)
    u = dist.Field(name='u', bases=basis)
    tau1 = dist.Field(name='tau1')
    tau2 = dist.Field(name='tau2')
    u_true = np.sin(x)
    f = dist.Field(bases=basis)
    f['g'] = -np.sin(x)
    dx = lambda A: d3.Differentiate(A, coord)
    lift = lambda A, n: d3.Lift(A, basis.derivative_basis(2), n)
    problem = d3.LBVP([u, tau1, tau2], namespace=locals())
    problem.add_equation("dx(dx(u)) + lift(tau1,-1) + lift(tau2,-2) = f")
    problem.add_equation("u(x='left') = 0")
    problem.add_equation("u(x='right') = 0")
    solver = problem.build_solver()
    solver.solve()
    assert np.allclose(u['g'], u_true)


def check_solve_jacobi_ncc(a0, b0, k_ncc, N=16, dealias=3/2):
    """dedalus/tests/test_cartesian_ncc.py:115-135 (k_arg = 0): solve f(x) u(x) = f(x) g(x) with a full-spectrum random coefficient f
    on the left-hand side (the NCC matrix of the coupled Jacobi axis) and the pseudospectral product on the right."""
    c = d3.Coordinate('x')
    d = d3.Distributor(c, dtype=dtype)
    b = d3.Jacobi(c, size=N, a=a0, b=b0, bounds=(0, 1), dealias=dealias)
    b_ncc = b.clone_with(a=a0+k_ncc, b=b0+k_ncc)
    f = d.Field(bases=b_ncc)
    g = d.Field(bases=b)
    u = d.Field(bases=b)
    f.fill_random('g')
    g.fill_random('g')
    problem = d3.LBVP([u])
    problem.add_equation((f*u, f*g))
    solver = problem.build_solver()
    solver.solve()
    assert np.allclose(u['c'], g['c'])
