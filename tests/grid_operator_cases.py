"""The reference's ufunc tests (dedalus/tests/test_grid_operators.py:35-85) restated against `dedalus_b200`, real dtype: every numpy /
scipy ufunc the reference registers for UnaryGridFunction, on Jacobi fields and on shell fields and operators."""
import numpy as np
import scipy.special as scp
import dedalus_b200 as d3

ufuncs = [np.absolute, np.sign, np.exp, np.exp2, np.log, np.log2, np.log10, np.sqrt, np.square, np.sin, np.cos, np.tan, np.arcsin,
          np.arccos, np.arctan, np.sinh, np.cosh, np.tanh, np.arcsinh, np.arccosh, np.arctanh, scp.erf]       # operators.py:534-556
dtype = np.float64


def check_jacobi_ufunc_field(a, b, N=16, dealias=1):
    """test_grid_operators.py:35-51"""
    for func in ufuncs:
        c = d3.Coordinate('x')
        d = d3.Distributor(c, dtype=dtype)
        basis = d3.Jacobi(c, size=N, a=a, b=b, bounds=(0, 1), dealias=dealias)
        x = d.local_grid(basis, scale=1)
        f = d.Field(bases=basis)
        f['g'] = 1 + x**2 if func is np.arccosh else x**2
        with np.errstate(all='ignore'):
            assert np.allclose(func(f)['g'], func(f['g'])), func.__name__


def check_shell_ufuncs(N=8, dealias=1):
    """test_grid_operators.py:54-85 with the shell basis"""
    c = d3.SphericalCoordinates('phi', 'theta', 'r')
    d = d3.Distributor(c, dtype=dtype)
    b = d3.ShellBasis(c, (2*N, N, N), radii=(0.5, 1), dtype=dtype, dealias=dealias)
    phi, theta, r = d.local_grids(b)
    for func in ufuncs:
        f = d.Field(bases=b)
        if func is np.arccosh:
            f['g'] = 1 + r**2
            a = 2
        else:
            f['g'] = r**2
            a = 0.5
        with np.errstate(all='ignore'):
            assert np.allclose(func(f)['g'], func(f['g'])), func.__name__
            assert np.allclose(func(a*f)['g'], func(a*f['g'])), func.__name__
