"""IMEX timesteppers (reference core/timesteppers.py): multistep (95-187, coefficient formulas 205-495)
and Runge-Kutta (552-644, tableaux 647-740).  The classes hold only the scheme definition; the device-resident
stage loop lives in dedalus_b200/solvers.py (one launch sequence per stage, no per-pencil Python loop).
"""
from collections import OrderedDict
import numpy as np

schemes = OrderedDict()


def add_scheme(scheme):
    schemes[scheme.__name__] = scheme
    return scheme


class MultistepIMEX:
    """a_j M.X(n-j) + b_j L.X(n-j) = c_j F(n-j)   (Wang & Ruuth 2008 coefficients, reference 34-63)."""
    stages = 1
    kind = "multistep"


class RungeKuttaIMEX:
    """(M + k H_ii L).X(n,i) = M.X(n,0) + k A_ij F(n,j) - k H_ij L.X(n,j)   (Ascher, Ruuth & Spiteri 1997)."""
    steps = 1
    kind = "rk"


@add_scheme
class CNAB1(MultistepIMEX):
    amax = bmax = cmax = 1
    steps = 1

    @classmethod
    def compute_coefficients(cls, timesteps, iteration):
        a = np.zeros(cls.amax + 1); b = np.zeros(cls.bmax + 1); c = np.zeros(cls.cmax + 1)
        k0 = timesteps[0]
        a[0] = 1 / k0; a[1] = -1 / k0
        b[0] = 1 / 2; b[1] = 1 / 2
        c[1] = 1
        return a, b, c


@add_scheme
class SBDF1(MultistepIMEX):
    amax = bmax = cmax = 1
    steps = 1

    @classmethod
    def compute_coefficients(cls, timesteps, iteration):
        a = np.zeros(cls.amax + 1); b = np.zeros(cls.bmax + 1); c = np.zeros(cls.cmax + 1)
        k0 = timesteps[0]
        a[0] = 1 / k0; a[1] = -1 / k0
        b[0] = 1
        c[1] = 1
        return a, b, c


@add_scheme
class CNAB2(MultistepIMEX):
    amax = bmax = cmax = 2
    steps = 2

    @classmethod
    def compute_coefficients(cls, timesteps, iteration):
        if iteration < 1:
            return CNAB1.compute_coefficients(timesteps, iteration)
        a = np.zeros(cls.amax + 1); b = np.zeros(cls.bmax + 1); c = np.zeros(cls.cmax + 1)
        k1, k0 = timesteps[0], timesteps[1]
        w1 = k1 / k0
        a[0] = 1 / k1; a[1] = -1 / k1
        b[0] = 1 / 2; b[1] = 1 / 2
        c[1] = 1 + w1 / 2; c[2] = -w1 / 2
        return a, b, c


@add_scheme
class MCNAB2(MultistepIMEX):
    amax = bmax = cmax = 2
    steps = 2

    @classmethod
    def compute_coefficients(cls, timesteps, iteration):
        if iteration < 1:
            return CNAB1.compute_coefficients(timesteps, iteration)
        a = np.zeros(cls.amax + 1); b = np.zeros(cls.bmax + 1); c = np.zeros(cls.cmax + 1)
        k1, k0 = timesteps[0], timesteps[1]
        w1 = k1 / k0
        a[0] = 1 / k1; a[1] = -1 / k1
        b[0] = (8 + 1 / w1) / 16; b[1] = (7 - 1 / w1) / 16; b[2] = 1 / 16
        c[1] = 1 + w1 / 2; c[2] = -w1 / 2
        return a, b, c


@add_scheme
class SBDF2(MultistepIMEX):
    amax = bmax = cmax = 2
    steps = 2

    @classmethod
    def compute_coefficients(cls, timesteps, iteration):
        if iteration < 1:
            return SBDF1.compute_coefficients(timesteps, iteration)
        a = np.zeros(cls.amax + 1); b = np.zeros(cls.bmax + 1); c = np.zeros(cls.cmax + 1)
        k1, k0 = timesteps[0], timesteps[1]
        w1 = k1 / k0
        a[0] = (1 + 2 * w1) / (1 + w1) / k1
        a[1] = -(1 + w1) / k1
        a[2] = w1**2 / (1 + w1) / k1
        b[0] = 1
        c[1] = 1 + w1; c[2] = -w1
        return a, b, c


@add_scheme
class CNLF2(MultistepIMEX):
    amax = bmax = cmax = 2
    steps = 2

    @classmethod
    def compute_coefficients(cls, timesteps, iteration):
        if iteration < 1:
            return CNAB1.compute_coefficients(timesteps, iteration)
        a = np.zeros(cls.amax + 1); b = np.zeros(cls.bmax + 1); c = np.zeros(cls.cmax + 1)
        k1, k0 = timesteps[0], timesteps[1]
        w1 = k1 / k0
        a[0] = 1 / (1 + w1) / k1
        a[1] = (w1 - 1) / k1
        a[2] = -w1**2 / (1 + w1) / k1
        b[0] = 1 / w1 / 2; b[1] = (1 - 1 / w1) / 2; b[2] = 1 / 2
        c[1] = 1
        return a, b, c


@add_scheme
class SBDF3(MultistepIMEX):
    amax = bmax = cmax = 3
    steps = 3

    @classmethod
    def compute_coefficients(cls, timesteps, iteration):
        if iteration < 2:
            return SBDF2.compute_coefficients(timesteps, iteration)
        a = np.zeros(cls.amax + 1); b = np.zeros(cls.bmax + 1); c = np.zeros(cls.cmax + 1)
        k2, k1, k0 = timesteps[0], timesteps[1], timesteps[2]
        w2 = k2 / k1; w1 = k1 / k0
        a[0] = (1 + w2 / (1 + w2) + w1 * w2 / (1 + w1 * (1 + w2))) / k2
        a[1] = (-1 - w2 - w1 * w2 * (1 + w2) / (1 + w1)) / k2
        a[2] = w2**2 * (w1 + 1 / (1 + w2)) / k2
        a[3] = -w1**3 * w2**2 * (1 + w2) / (1 + w1) / (1 + w1 + w1 * w2) / k2
        b[0] = 1
        c[1] = (1 + w2) * (1 + w1 * (1 + w2)) / (1 + w1)
        c[2] = -w2 * (1 + w1 * (1 + w2))
        c[3] = w1 * w1 * w2 * (1 + w2) / (1 + w1)
        return a, b, c


@add_scheme
class SBDF4(MultistepIMEX):
    amax = bmax = cmax = 4
    steps = 4

    @classmethod
    def compute_coefficients(cls, timesteps, iteration):
        if iteration < 3:
            return SBDF3.compute_coefficients(timesteps, iteration)
        a = np.zeros(cls.amax + 1); b = np.zeros(cls.bmax + 1); c = np.zeros(cls.cmax + 1)
        k3, k2, k1, k0 = timesteps[0], timesteps[1], timesteps[2], timesteps[3]
        w3 = k3 / k2; w2 = k2 / k1; w1 = k1 / k0
        A1 = 1 + w1 * (1 + w2)
        A2 = 1 + w2 * (1 + w3)
        A3 = 1 + w1 * A2
        a[0] = (1 + w3 / (1 + w3) + w2 * w3 / A2 + w1 * w2 * w3 / A3) / k3
        a[1] = (-1 - w3 * (1 + w2 * (1 + w3) / (1 + w2) * (1 + w1 * A2 / A1))) / k3
        a[2] = w3 * (w3 / (1 + w3) + w2 * w3 * (A3 + w1) / (1 + w1)) / k3
        a[3] = -w2**3 * w3**2 * (1 + w3) / (1 + w2) * A3 / A2 / k3
        a[4] = (1 + w3) / (1 + w1) * A2 / A1 * w1**4 * w2**3 * w3**2 / A3 / k3
        b[0] = 1
        c[1] = w2 * (1 + w3) / (1 + w2) * ((1 + w3) * (A3 + w1) + (1 + w1) / w2) / A1
        c[2] = -A2 * A3 * w3 / (1 + w1)
        c[3] = w2**2 * w3 * (1 + w3) / (1 + w2) * A3
        c[4] = -w1**3 * w2**2 * w3 * (1 + w3) / (1 + w1) * A2 / A1
        return a, b, c


@add_scheme
class RK111(RungeKuttaIMEX):
    stages = 1
    c = np.array([0, 1])
    A = np.array([[0, 0], [1, 0]])
    H = np.array([[0, 0], [0, 1]])


@add_scheme
class RK222(RungeKuttaIMEX):
    stages = 2
    _g = (2 - np.sqrt(2)) / 2
    _d = 1 - 1 / _g / 2
    c = np.array([0, _g, 1])
    A = np.array([[0, 0, 0], [_g, 0, 0], [_d, 1 - _d, 0]])
    H = np.array([[0, 0, 0], [0, _g, 0], [0, 1 - _g, _g]])


@add_scheme
class RK443(RungeKuttaIMEX):
    stages = 4
    c = np.array([0, 1/2, 2/3, 1/2, 1])
    A = np.array([[0, 0, 0, 0, 0],
                  [1/2, 0, 0, 0, 0],
                  [11/18, 1/18, 0, 0, 0],
                  [5/6, -5/6, 1/2, 0, 0],
                  [1/4, 7/4, 3/4, -7/4, 0]])
    H = np.array([[0, 0, 0, 0, 0],
                  [0, 1/2, 0, 0, 0],
                  [0, 1/6, 1/2, 0, 0],
                  [0, -1/2, 1/2, 1/2, 0],
                  [0, 3/2, -3/2, 1/2, 1/2]])


@add_scheme
class RKSMR(RungeKuttaIMEX):
    stages = 3
    _a1, _a2, _a3 = (29/96, -3/40, 1/6)
    _b1, _b2, _b3 = (37/160, 5/24, 1/6)
    _g1, _g2, _g3 = (8/15, 5/12, 3/4)
    _z2, _z3 = (-17/60, -5/12)
    c = np.array([0, 8/15, 2/3, 1])
    A = np.array([[0, 0, 0, 0],
                  [_g1, 0, 0, 0],
                  [_g1 + _z2, _g2, 0, 0],
                  [_g1 + _z2, _g2 + _z3, _g3, 0]])
    H = np.array([[0, 0, 0, 0],
                  [_a1, _b1, 0, 0],
                  [_a1, _b1 + _a2, _b2, 0],
                  [_a1, _b1 + _a2, _b2 + _a3, _b3]])


class RKGFY(RungeKuttaIMEX):
    stages = 2
    c = np.array([0, 1, 1])
    A = np.array([[0, 0, 0], [1, 0, 0], [0.5, 0.5, 0]])
    H = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5]])
