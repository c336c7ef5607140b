"""IMEX scheme tables restated for the oracle (TEST INFRASTRUCTURE).  Reference: core/timesteppers.py:647-740
(Runge-Kutta tableaux) and 205-495 (multistep coefficients, Wang & Ruuth 2008)."""
import numpy as np

_g = (2 - np.sqrt(2)) / 2
_d = 1 - 1 / _g / 2
RK = {
    "RK111": dict(c=[0, 1], A=[[0, 0], [1, 0]], H=[[0, 0], [0, 1]]),
    "RK222": dict(c=[0, _g, 1], A=[[0, 0, 0], [_g, 0, 0], [_d, 1 - _d, 0]], H=[[0, 0, 0], [0, _g, 0], [0, 1 - _g, _g]]),
    "RK443": dict(c=[0, 1/2, 2/3, 1/2, 1],
                  A=[[0, 0, 0, 0, 0], [1/2, 0, 0, 0, 0], [11/18, 1/18, 0, 0, 0], [5/6, -5/6, 1/2, 0, 0], [1/4, 7/4, 3/4, -7/4, 0]],
                  H=[[0, 0, 0, 0, 0], [0, 1/2, 0, 0, 0], [0, 1/6, 1/2, 0, 0], [0, -1/2, 1/2, 1/2, 0], [0, 3/2, -3/2, 1/2, 1/2]]),
}
for v in RK.values():
    for k in v:
        v[k] = np.array(v[k], dtype=float)


def sbdf1(k0):
    return np.array([1 / k0, -1 / k0]), np.array([1.0, 0.0]), np.array([0.0, 1.0])


def sbdf2(k1, k0, iteration):
    if iteration < 1:
        a, b, c = sbdf1(k1)
        return np.append(a, 0), np.append(b, 0), np.append(c, 0)
    w1 = k1 / k0
    a = np.array([(1 + 2 * w1) / (1 + w1) / k1, -(1 + w1) / k1, w1**2 / (1 + w1) / k1])
    b = np.array([1.0, 0.0, 0.0])
    c = np.array([0.0, 1 + w1, -w1])
    return a, b, c
