"""Linear boundary value problems modelled on the Cartesian cases of the reference's dedalus/tests/test_lbvp.py:38-111 and
test_cartesian_ncc.py:115-135 (real dtype).  Every case has an analytic solution; the assertion is numpy's default allclose, as in the
reference."""
import numpy as np
import dedalus_b200 as d3

REAL = np.float64


def _line(basis_factory):
    coord = d3.Coordinate('x')
    dist = d3.Distributor(coord, dtype=REAL)
    basis = basis_factory(coord)
    return coord, dist, basis, dist.local_grid(basis)


def check_algebraic():
    """test_lbvp.py:38-55 -- no basis at all: v u = F with constants"""
    coord = d3.Coordinate('x')
    dist = d3.Distributor(coord, dtype=REAL)
    unknown, coef, rhs = dist.Field(name='u'), dist.Field(name='v'), dist.Field(name='F')
    coef['g'], rhs['g'] = -1, -3
    bvp = d3.LBVP([unknown], namespace=dict(u=unknown, v=coef, F=rhs))
    bvp.add_equation("v*u = F")
    bvp.build_solver().solve()
    assert np.allclose(unknown['g'], 3)


def check_poisson_fourier(N=32):
    """test_lbvp.py:58-82 -- u'' + c = -sin x on a periodic interval, with the gauge constant c and a zero-mean condition"""
    coord, dist, basis, x = _line(lambda c: d3.Fourier(c, size=N, bounds=(0, 2 * np.pi), dtype=REAL))
    u, gauge, forcing = dist.Field(name='u', bases=basis), dist.Field(name='c'), dist.Field(bases=basis)
    forcing['g'] = -np.sin(x)
    ns = dict(u=u, g=gauge, f=forcing, dx=lambda A: d3.Differentiate(A, coord), integ=lambda A: d3.Integrate(A, coord))
    bvp = d3.LBVP([u, gauge], namespace=ns)
    bvp.add_equation("dx(dx(u)) + g = f")
    bvp.add_equation("integ(u) = 0")
    bvp.build_solver().solve()
    assert np.allclose(u['g'], np.sin(x))


def check_poisson_jacobi(a, b, N=32):
    """test_lbvp.py:85-111 -- u'' = -sin x with homogeneous Dirichlet data on (0, 2 pi), two tau terms, Chebyshev or Legendre"""
    coord, dist, basis, x = _line(lambda c: d3.Jacobi(c, size=N, bounds=(0, 2 * np.pi), a=a, b=b))
    u, t1, t2, forcing = dist.Field(name='u', bases=basis), dist.Field(name='tau1'), dist.Field(name='tau2'), dist.Field(bases=basis)
    forcing['g'] = -np.sin(x)
    ns = dict(u=u, tau1=t1, tau2=t2, f=forcing, dx=lambda A: d3.Differentiate(A, coord),
              lift=lambda A, n: d3.Lift(A, basis.derivative_basis(2), n))
    bvp = d3.LBVP([u, t1, t2], namespace=ns)
    bvp.add_equation("dx(dx(u)) + lift(tau1,-1) + lift(tau2,-2) = f")
    bvp.add_equation("u(x='left') = 0")
    bvp.add_equation("u(x='right') = 0")
    bvp.build_solver().solve()
    assert np.allclose(u['g'], np.sin(x))


def check_solve_jacobi_ncc(a0, b0, k_ncc, N=16, dealias=3/2):
    """test_cartesian_ncc.py:115-135 (k_arg = 0) -- f(x) u(x) = f(x) g(x) with a full-spectrum random coefficient on the left (the NCC
    matrix of the coupled Jacobi axis) and the pseudospectral product on the right; u must come back as g."""
    coord, dist, basis, _ = _line(lambda c: d3.Jacobi(c, size=N, a=a0, b=b0, bounds=(0, 1), dealias=dealias))
    coef = dist.Field(bases=basis.clone_with(a=a0 + k_ncc, b=b0 + k_ncc))
    target, unknown = dist.Field(bases=basis), dist.Field(bases=basis)
    coef.fill_random('g')
    target.fill_random('g')
    bvp = d3.LBVP([unknown])
    bvp.add_equation((coef * unknown, coef * target))
    bvp.build_solver().solve()
    assert np.allclose(unknown['c'], target['c'])
