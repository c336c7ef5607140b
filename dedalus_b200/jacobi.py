"""Jacobi-polynomial algebra for unit-weight-normalised polynomials on [-1, 1] (host, setup-time only).

Own derivation from the classical identities (Abramowitz & Stegun 22.7/22.8; DLMF 18.9):
    P_n^(a,b)   = [(n+a+b+1) P_n^(a+1,b) - (n+b) P_{n-1}^(a+1,b)] / (2n+a+b+1)
    P_n^(a,b)   = [(n+a+b+1) P_n^(a,b+1) + (n+a) P_{n-1}^(a,b+1)] / (2n+a+b+1)
    d/dz P_n^(a,b) = (n+a+b+1)/2 * P_{n-1}^(a+1,b+1)
    h_n^(a,b)   = 2^(a+b+1) G(n+a+1) G(n+b+1) / ((2n+a+b+1) G(n+a+b+1) n!)
with p_n = P_n / sqrt(h_n), so that  int (1-z)^a (1+z)^b p_m p_n dz = delta_mn.

Plays the role of the reference's tools/jacobi.py:203-260 + libraries/dedalus_sphere/jacobi.py
(conversion/differentiation/Jacobi matrices, Gauss grids, polynomial evaluation); parity is pinned by
tests/test_jacobi.py against tests/golden/transforms.npz (jop_* arrays dumped from the reference).
"""
import numpy as np
from scipy import sparse
from scipy.special import gammaln, roots_jacobi

LD = np.longdouble


def mass(a, b):
    """int_{-1}^{1} (1-z)^a (1+z)^b dz."""
    return float(np.exp((a + b + 1) * np.log(2.0) + gammaln(a + 1) + gammaln(b + 1) - gammaln(a + b + 2)))


def log_norm(n, a, b):
    """log h_n^(a,b) for integer array n >= 0."""
    n = np.asarray(n, dtype=float)
    out = ((a + b + 1) * np.log(2.0) + gammaln(n + a + 1) + gammaln(n + b + 1)
           - gammaln(n + a + b + 1) - gammaln(n + 1) - np.log(np.where(2*n + a + b + 1 == 0, 1.0, 2*n + a + b + 1)))
    # n = 0 with a+b+1 = 0: Gamma(0)*0 -> use the integral directly
    h0 = np.log(mass(a, b))
    return np.where(n == 0, h0, out)


def _raise_one(N, a, b, which):
    """Matrix C with p^(a,b) coefficients -> p^(a+1,b) (which='a') or p^(a,b+1) (which='b') coefficients."""
    n = np.arange(N, dtype=float)
    if which == 'a':
        a1, b1 = a + 1, b
        off = -(n + b)
    else:
        a1, b1 = a, b + 1
        off = (n + a)
    den = 2*n + a + b + 1
    with np.errstate(divide='ignore', invalid='ignore'):
        diag = (n + a + b + 1) / den
        sup = off / den
    if den[0] == 0:          # a+b = -1, n = 0: P_0 = P_0
        diag[0] = 1.0
    sup[0] = 0.0
    ln0 = log_norm(n, a, b)
    ln1 = log_norm(n, a1, b1)
    d = diag * np.exp(0.5 * (ln1 - ln0))
    # entry (n-1, n): coefficient of p_{n-1}^(new) in p_n^(old)
    s = np.zeros(N)
    s[1:] = sup[1:] * np.exp(0.5 * (ln1[:-1] - ln0[1:]))
    return sparse.diags([d, s[1:]], [0, 1], shape=(N, N), format='csr')


def conversion_matrix(N, a0, b0, a1, b1):
    """Coefficients in p^(a0,b0) -> coefficients in p^(a1,b1) (integer raises); upper triangular, bandwidth da+db."""
    da, db = a1 - a0, b1 - b0
    if abs(da - round(da)) > 1e-12 or abs(db - round(db)) > 1e-12 or da < -1e-12 or db < -1e-12:
        raise ValueError("Jacobi conversion requires non-negative integer parameter increments.")
    C = sparse.identity(N, format='csr')
    a, b = a0, b0
    for _ in range(int(round(db))):
        C = _raise_one(N, a, b, 'b') @ C
        b += 1
    for _ in range(int(round(da))):
        C = _raise_one(N, a, b, 'a') @ C
        a += 1
    return C.tocsr()


def differentiation_matrix(N, a, b):
    """d/dz: p^(a,b) coefficients -> p^(a+1,b+1) coefficients (square N x N, single superdiagonal)."""
    n = np.arange(N, dtype=float)
    ln0 = log_norm(n, a, b)
    ln1 = log_norm(n, a + 1, b + 1)
    s = np.zeros(N)
    s[1:] = 0.5 * (n[1:] + a + b + 1) * np.exp(0.5 * (ln1[:-1] - ln0[1:]))
    return sparse.diags([s[1:]], [1], shape=(N, N), format='csr')


def polynomials(N, a, b, z, half_log_weight=False):
    """Unit-normalised p_n^(a,b)(z), n < N, as an (N, len(z)) float64 array (long-double recurrence).
    half_log_weight=True multiplies by sqrt((1-z)^a (1+z)^b), applied in log space inside the long-double scaling: for large
    a, b (spherical harmonics of high order) the envelope underflows and the polynomial values overflow double range
    separately, while their product is O(1) (the reference starts its recurrence from the envelope for the same reason,
    libraries/dedalus_sphere/sphere.py:59-64)."""
    z = np.atleast_1d(np.asarray(z, dtype=LD))
    P = np.zeros((max(N, 2), z.size), dtype=LD)
    P[0] = 1
    P[1] = ((a + b + 2) * z + (a - b)) / 2
    for n in range(1, N - 1):
        c = 2*n + a + b
        a1 = 2 * (n + 1) * (n + a + b + 1) * c
        a2 = (c + 1) * (a*a - b*b)
        a3 = c * (c + 1) * (c + 2)
        a4 = 2 * (n + a) * (n + b) * (c + 2)
        P[n + 1] = ((a2 + a3 * z) * P[n] - a4 * P[n - 1]) / a1
    ln = log_norm(np.arange(max(N, 2)), a, b)
    if half_log_weight:
        zl = z.astype(LD)
        with np.errstate(divide='ignore'):
            lw = 0.5 * (a * np.log1p(-zl) + b * np.log1p(zl))
        P = P * np.exp(-0.5 * ln.astype(LD)[:, None] + lw[None, :])
    else:
        P = P * np.exp(-0.5 * ln).astype(LD)[:, None]
    return np.asarray(P[:N], dtype=np.float64)


def gauss_grid(N, a, b):
    """Gauss-Jacobi nodes (ascending) and weights for weight (1-z)^a (1+z)^b."""
    if a == b == -0.5:
        j = np.arange(N, dtype=LD)
        z = -np.cos(np.pi * (2*j + 1) / (2 * LD(N)))
        w = np.full(N, np.pi / N)
        return np.asarray(z, dtype=np.float64), w
    z, w = roots_jacobi(N, a, b)
    return z, w


def jacobi_matrix(N, a, b):
    """Multiplication by z in the p^(a,b) basis (symmetric tridiagonal, N x N)."""
    n = np.arange(N + 1, dtype=float)
    c = 2*n + a + b
    with np.errstate(divide='ignore', invalid='ignore'):
        d = (b*b - a*a) / (c * (c + 2))
    if c[0] == 0 or c[0] + 2 == 0:
        d[0] = (b - a) / (a + b + 2)
    # classical: z P_n = A_n P_{n+1} + B_n P_n + C_n P_{n-1}, A_n = 2(n+1)(n+a+b+1)/((c+1)(c+2))
    with np.errstate(divide='ignore', invalid='ignore'):
        A = 2 * (n + 1) * (n + a + b + 1) / ((c + 1) * (c + 2))
    if c[0] + 1 == 0:
        A[0] = 2 / (a + b + 2)
    ln = log_norm(n, a, b)
    off = A[:-1] * np.exp(0.5 * (ln[1:] - ln[:-1]))     # <p_{n+1}| z |p_n>
    return sparse.diags([off[:N-1], d[:N], off[:N-1]], [-1, 0, 1], shape=(N, N), format='csr')


def integration_vector(N, a, b):
    """int_{-1}^{1} p_n^(a,b)(z) dz for n < N (Gauss-Legendre, exact)."""
    zl, wl = np.polynomial.legendre.leggauss(N + 1)
    return polynomials(N, a, b, zl) @ wl


def interpolation_vector(N, a, b, z):
    """p_n^(a,b)(z) for n < N at one native position z."""
    return polynomials(N, a, b, np.array([z]))[:, 0]
