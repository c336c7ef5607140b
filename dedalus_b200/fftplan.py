"""Host-side plan tables for csrc/fft.cu: radix factorisation, twiddles and the digit-reversal permutation of
the in-place mixed-radix FFT (DIF forward / DIT backward).  Plays the role of FFTW plan creation
(libraries/fftw/fftw_wrappers.pyx:61-214, 217-325) -- but a plan here is just four small tables."""
import numpy as np

MAX_RADIX = 16


def factorize(n):
    """Radices in DIF order (4s first).  Returns None if n has a prime factor > MAX_RADIX."""
    f = []
    while n % 4 == 0:
        f.append(4); n //= 4
    for r in (2, 3, 5, 7, 11, 13):
        while n % r == 0:
            f.append(r); n //= r
    if n != 1:
        return None
    return f


def digit_reversal(n, radices):
    """perm[p] = frequency index held at position p after the DIF passes; iperm = inverse."""
    pos = np.arange(n)
    freq = np.zeros(n, dtype=np.int64)
    L, mult, rem = n, 1, pos.copy()
    for r in radices:
        m = L // r
        d = rem // m
        rem = rem % m
        freq += d * mult
        mult *= r
        L = m
    iperm = np.zeros(n, dtype=np.int64)
    iperm[freq] = pos
    return freq.astype(np.int32), iperm.astype(np.int32)


def _cis(num, den):
    """exp(-2 pi i num/den) evaluated in long double, returned as (len, 2) float64."""
    x = np.asarray(num, dtype=np.longdouble) / np.longdouble(den)
    ang = -2 * np.pi * x.astype(np.longdouble)
    return np.stack([np.cos(ang), np.sin(ang)], axis=-1).astype(np.float64)


class HostPlan:
    """Tables for a length-n transform.  kind: 'real' (RealFourier / Chebyshev) or 'complex'."""

    def __init__(self, n, kind):
        self.n = int(n)
        self.kind = kind
        if kind == 'real' and n % 2 == 0 and n >= 2 and factorize(n // 2) is not None:
            self.half, self.nc = 1, n // 2
        else:
            self.half, self.nc = 0, n
        rad = factorize(self.nc)
        if rad is None:
            # a single generic pass is only possible for nc <= MAX_RADIX
            raise NotImplementedError(
                f"transform length {n}: prime factor > {MAX_RADIX} is not supported by the shared-memory FFT")
        if len(rad) > 16:
            raise NotImplementedError("too many radix passes")
        self.radices = rad
        self.perm, self.iperm = digit_reversal(self.nc, rad)
        self.tw = _cis(np.arange(self.nc), self.nc)                       # exp(-2 pi i j / nc)
        self.twr = _cis(np.arange(self.nc + 1), self.n)                   # exp(-2 pi i k / n)
        self.twq = _cis(np.arange(self.n), 4 * self.n)                    # exp(-i pi k / (2 n))
        self.twn = _cis(np.arange(self.n), self.n)                        # exp(-2 pi i j / n)
