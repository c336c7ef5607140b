"""Numpy/scipy restatement of the reference's 1-D transforms (TEST INFRASTRUCTURE, see oracle/__init__.py).

Each function cites the reference code it follows (paths under /root/reference/dedalus).  Two flavours:
matrix transforms (the reference's ground-truth MMT classes) and fast transforms (its scipy-FFT classes).
"""
import functools
import numpy as np
import scipy.fft
import scipy.linalg
from scipy import sparse
from scipy.special import eval_jacobi, gammaln, roots_jacobi


def axslice(axis, start, stop, step=None):
    return (slice(None),) * axis + (slice(start, stop, step),)


def apply_along(mat, data, axis):
    """core/transforms.py:57-63 SeparableMatrixTransform -> tools/array.py:104-129 apply_dense."""
    return np.moveaxis(np.tensordot(mat, data, axes=(1, axis)), 0, axis)


# ---- Real Fourier ------------------------------------------------------------------------------------------
def rf_matrices(N, M):
    """core/transforms.py:387-424 RealFourierMMT.forward_matrix / backward_matrix."""
    KN, KM = (N - 1) // 2, (M - 1) // 2
    Kmax = min(KN, KM)
    Mm = max(2, M)
    wav = np.repeat(np.arange(KM + 1), 2)[:Mm]
    K = wav[::2, None]; X = np.arange(N)[None, :]; dX = N / 2 / np.pi
    fwd = np.zeros((Mm, N))
    fwd[0::2] = (2 / N) * np.cos(K * X / dX)
    fwd[1::2] = -(2 / N) * np.sin(K * X / dX)
    fwd[0] = 1 / N
    fwd *= (wav[:, None] <= Kmax)
    bwd = np.zeros((N, Mm))
    bwd[:, 0::2] = np.cos(K.T * X.T / dX)
    bwd[:, 1::2] = -np.sin(K.T * X.T / dX)
    bwd *= (wav[None, :] <= Kmax)
    return fwd, bwd


def rf_forward_fft(g, M, axis):
    """core/transforms.py:516-521 ScipyRealFFT.forward + 472-487 unpack_rescale."""
    N = g.shape[axis]
    Kmax = min((N - 1) // 2, (M - 1) // 2)
    temp = scipy.fft.rfft(g, axis=axis)
    shp = list(g.shape); shp[axis] = M
    c = np.zeros(shp)
    c[axslice(axis, 0, 1)] = temp[axslice(axis, 0, 1)].real / N
    pos = temp[axslice(axis, 1, Kmax + 1)]
    c[axslice(axis, 2, 2 * (Kmax + 1), 2)] = pos.real * (2 / N)
    c[axslice(axis, 3, 2 * (Kmax + 1), 2)] = pos.imag * (2 / N)
    return c


def rf_backward_fft(c, N, axis):
    """core/transforms.py:523-534 ScipyRealFFT.backward + 489-509 repack_rescale."""
    M = c.shape[axis]
    Kmax = min((N - 1) // 2, (M - 1) // 2)
    shp = list(c.shape); shp[axis] = N // 2 + 1
    temp = np.zeros(shp, dtype=np.complex128)
    temp[axslice(axis, 0, 1)] = c[axslice(axis, 0, 1)] * N
    pos = temp[axslice(axis, 1, Kmax + 1)]
    pos.real[...] = c[axslice(axis, 2, 2 * (Kmax + 1), 2)] * (N / 2)
    pos.imag[...] = c[axslice(axis, 3, 2 * (Kmax + 1), 2)] * (N / 2)
    return scipy.fft.irfft(temp, axis=axis, n=N)


# ---- Complex Fourier ---------------------------------------------------------------------------------------
def cf_matrices(N, M):
    """core/transforms.py:211-237 ComplexFourierMMT."""
    KM = (M - 1) // 2
    Kmax = min((N - 1) // 2, KM)
    wav = (np.arange(M) + KM) % M - KM
    X = np.arange(N); dX = N / 2 / np.pi
    fwd = np.exp(-1j * wav[:, None] * X[None, :] / dX) / N * (np.abs(wav[:, None]) <= Kmax)
    bwd = np.exp(1j * wav[None, :] * X[:, None] / dX) * (np.abs(wav[None, :]) <= Kmax)
    return fwd, bwd


# ---- Jacobi / Chebyshev ------------------------------------------------------------------------------------
def log_norm(n, a, b):
    n = np.asarray(n, dtype=float)
    return ((a + b + 1) * np.log(2.0) + gammaln(n + a + 1) + gammaln(n + b + 1)
            - gammaln(n + a + b + 1) - gammaln(n + 1) - np.log(2 * n + a + b + 1))


def jacobi_polynomials(M, a, b, z):
    """Unit-normalised polynomials, libraries/dedalus_sphere/jacobi.py:30-81 (there by recurrence; here through
    scipy.special.eval_jacobi and the closed-form norm; n=0 with a+b=-1 handled by the Beta integral)."""
    z = np.asarray(z, dtype=float)
    n = np.arange(M)
    P = np.stack([eval_jacobi(k, a, b, z) for k in n], axis=0)
    with np.errstate(divide='ignore', invalid='ignore'):
        ln = log_norm(n, a, b)
    ln[0] = (a + b + 1) * np.log(2.0) + gammaln(a + 1) + gammaln(b + 1) - gammaln(a + b + 2)
    return P * np.exp(-0.5 * ln)[:, None]


def jacobi_grid(N, a0, b0):
    """Gauss-Jacobi nodes / weights, libraries/dedalus_sphere/jacobi.py:83-144 (Chebyshev closed form 110-111, 131-138)."""
    if a0 == b0 == -0.5:
        j = np.arange(N)
        return -np.cos(np.pi * (2 * j + 1) / (2 * N)), np.full(N, np.pi / N)
    return roots_jacobi(N, a0, b0)


def _quad(a, b, K):
    """Gauss quadrature exact for weight (1-z)^a (1+z)^b times polynomials of degree < 2K."""
    return roots_jacobi(K, a, b)


@functools.lru_cache(maxsize=64)
def _conversion_cached(N, a0, b0, a1, b1):
    C = jacobi_conversion(N, a0, b0, a1, b1)
    C[np.abs(C) < 1e-13] = 0
    nd = int(round((a1 - a0) + (b1 - b0))) + 1
    ab = np.zeros((nd, N))                       # LAPACK upper-band storage for solve_banded((0, nd-1), ...)
    for d in range(nd):
        ab[nd - 1 - d, d:] = np.diagonal(C, d)
    return sparse.csr_matrix(C), ab, nd


def _apply_sparse_along(Cs, data, axis):
    """tools/array.py:171-203 apply_sparse along an axis."""
    moved = np.moveaxis(data, axis, 0)
    out = Cs @ moved.reshape(moved.shape[0], -1)
    return np.moveaxis(out.reshape((Cs.shape[0],) + moved.shape[1:]), 0, axis)


def jacobi_conversion(N, a0, b0, a1, b1):
    """tools/jacobi.py:229-245 conversion_matrix, computed here by exact Gauss quadrature projection
    C_ij = int w1 p_i^(a1,b1) p_j^(a0,b0)."""
    z, w = _quad(a1, b1, N + 2)
    return (jacobi_polynomials(N, a1, b1, z) * w) @ jacobi_polynomials(N, a0, b0, z).T


def jacobi_differentiation(N, a, b):
    """tools/jacobi.py:247-248 differentiation_matrix: D_ij = int w(a+1,b+1) p_i^(a+1,b+1) d/dz p_j^(a,b),
    with d/dz P_n^(a,b) = (n+a+b+1)/2 P_{n-1}^(a+1,b+1)."""
    z, w = _quad(a + 1, b + 1, N + 2)
    n = np.arange(N)
    dP = np.zeros((N, z.size))
    for k in range(1, N):
        dP[k] = 0.5 * (k + a + b + 1) * eval_jacobi(k - 1, a + 1, b + 1, z)
    with np.errstate(divide='ignore', invalid='ignore'):
        ln = log_norm(n, a, b)
    ln[0] = (a + b + 1) * np.log(2.0) + gammaln(a + 1) + gammaln(b + 1) - gammaln(a + b + 2)
    dp = dP * np.exp(-0.5 * ln)[:, None]
    return (jacobi_polynomials(N, a + 1, b + 1, z) * w) @ dp.T


def jacobi_integration(N, a, b):
    """tools/jacobi.py:253-260 integration_vector (Legendre quadrature)."""
    zl, wl = np.polynomial.legendre.leggauss(N + 1)
    return jacobi_polynomials(N, a, b, zl) @ wl


def jacobi_matrices(N, M, a, b, a0, b0):
    """core/transforms.py:114-158 JacobiMMT.forward_matrix / backward_matrix (DEALIAS_BEFORE_CONVERTING=True)."""
    grid, wts = jacobi_grid(N, a0, b0)
    base = jacobi_polynomials(max(M, N), a0, b0, grid) * wts
    base[N:, :] = 0
    base = base[:M]
    fwd = base if (a, b) == (a0, b0) else jacobi_conversion(base.shape[0], a0, b0, a, b) @ base
    poly = jacobi_polynomials(M, a, b, grid)
    poly[N:, :] = 0
    return fwd[:M], poly.T


def cheb_forward_fft(g, M, axis, a=-0.5, b=-0.5):
    """core/transforms.py:749-756 ScipyDCT.forward + 715-746 rescale + 801-874 FastChebyshevTransform forward."""
    N = g.shape[axis]
    temp = scipy.fft.dct(g, type=2, axis=axis)
    convert = (a, b) != (-0.5, -0.5)
    Md = N if convert else M
    Kmax = min(N - 1, Md - 1)
    shp = list(g.shape); shp[axis] = Md
    out = np.zeros(shp)
    out[axslice(axis, 0, 1)] = temp[axslice(axis, 0, 1)] * (np.sqrt(np.pi) / N / 2)
    out[axslice(axis, 1, Kmax + 1)] = temp[axslice(axis, 1, Kmax + 1)] * (np.sqrt(np.pi / 2) / N)
    out[axslice(axis, 1, Kmax + 1, 2)] *= -1
    if not convert:
        return out
    Kin = min(M, N)
    Cs = _conversion_cached(max(M, N), -0.5, -0.5, a, b)[0][:M, :Kin]
    return _apply_sparse_along(Cs, out[axslice(axis, 0, Kin)], axis)


def cheb_backward_fft(c, N, axis, a=-0.5, b=-0.5):
    """core/transforms.py:758-768 ScipyDCT.backward + 876-890 FastChebyshevTransform backward (incl. the
    solve_upper_sparse back-substitution of the conversion, tools/linalg.pyx:20-82)."""
    M = c.shape[axis]
    Kmax = min(N - 1, M - 1)
    data = np.array(c, dtype=float, copy=True)
    if M > N:
        data[axslice(axis, Kmax + 1, None)] = 0
    if (a, b) != (-0.5, -0.5):
        _, ab, nd = _conversion_cached(M, -0.5, -0.5, a, b)
        moved = np.moveaxis(data, axis, 0)
        sol = scipy.linalg.solve_banded((0, nd - 1), ab, moved.reshape(M, -1))      # upper-triangular banded back-substitution
        data = np.moveaxis(sol.reshape(moved.shape), 0, axis).copy()
    data[axslice(axis, 1, Kmax + 1, 2)] *= -1
    shp = list(c.shape); shp[axis] = N
    temp = np.zeros(shp)
    temp[axslice(axis, 0, 1)] = data[axslice(axis, 0, 1)] / np.sqrt(np.pi)
    temp[axslice(axis, 1, Kmax + 1)] = data[axslice(axis, 1, Kmax + 1)] * (0.5 / np.sqrt(np.pi / 2))
    return scipy.fft.dct(temp, type=3, axis=axis)


# ---------------------------------------------------------------------------------------------------------
# Spin-weighted spherical harmonic colatitude transform (reference core/transforms.py:1251-1340).
# Harmonics restated from libraries/dedalus_sphere/sphere.py:43-64 + jacobi.py:28-80, 147-172:
#   Y_{l,m,s}(z) = (-1)^max(m,-s) sqrt((1-z)^a (1+z)^b) p_k^(a,b)(z),  a = |m+s|, b = |m-s|, k = l - max(|m|,|s|),
# p_k the unit-weight-normalised Jacobi polynomials; Gauss-Legendre quadrature in z = cos(theta) with Ntheta nodes.
# ---------------------------------------------------------------------------------------------------------
def swsh_quadrature(Ntheta):
    z, w = roots_jacobi(Ntheta, 0.0, 0.0)
    return z, w


def swsh_harmonics(Lmax, m, s, z):
    n = Lmax + 1 - max(abs(m), abs(s))
    a, b = abs(m + s), abs(m - s)
    if n <= 0:
        return np.zeros((0, len(z)))
    # envelope and normalisation combined in log space (long double): separately they under- / overflow for large m
    zl = np.asarray(z, dtype=np.longdouble)
    k = np.arange(n)
    with np.errstate(divide='ignore', invalid='ignore'):
        ln = log_norm(k, a, b)
    ln[0] = (a + b + 1) * np.log(2.0) + gammaln(a + 1) + gammaln(b + 1) - gammaln(a + b + 2)
    P = np.stack([eval_jacobi(int(q), a, b, z) for q in k], axis=0).astype(np.longdouble)
    scale = np.exp(-0.5 * ln.astype(np.longdouble)[:, None] + 0.5 * (a * np.log1p(-zl) + b * np.log1p(zl))[None, :])
    return np.asarray(P * scale, dtype=float) * ((-1.0) ** max(m, -s))


@functools.lru_cache(maxsize=None)
def swsh_matrices(Ntheta, Lmax, m, s):
    """(forward (Lmax+1-|m|, Ntheta), backward (Ntheta, Lmax+1-|m|)) padded with zero rows for l < |s| and zeroed for
    l >= Ntheta (core/transforms.py:1291-1340)."""
    z, w = swsh_quadrature(Ntheta)
    Y = swsh_harmonics(Lmax, m, s, z)
    Lmin = max(abs(m), abs(s))
    F = np.zeros((Lmax + 1 - abs(m), Ntheta)); B = np.zeros((Ntheta, Lmax + 1 - abs(m)))
    F[Lmin - abs(m):, :] = Y * w[None, :]
    B[:, Lmin - abs(m):] = Y.T
    F[max(Ntheta - abs(m), 0):, :] = 0
    B[:, max(Ntheta - abs(m), 0):] = 0
    return F, B


def _ell_index(row, nell):
    start, stop, step = int(row[5]), int(row[6]), int(row[7])
    return np.arange(nell)[slice(start, None if stop < 0 else stop, step)]


def swsh_forward(g, c, m_maps, Ntheta, Lmax, s):
    """g: (..., Nphi_cg, Ntheta) -> c: (..., Nphi_c, Nell) in place; m_maps rows (m, mg0, mg1, mc0, mc1, ell start/stop/step)."""
    for row in np.asarray(m_maps):
        m = int(row[0])
        if abs(m) > Lmax:
            continue
        F, _ = swsh_matrices(Ntheta, Lmax, m, s)
        ell = _ell_index(row, c.shape[-1])
        c[..., row[3]:row[4], ell] = np.einsum('kt,...it->...ik', F, g[..., row[1]:row[2], :])
    return c


def swsh_backward(c, g, m_maps, Ntheta, Lmax, s):
    for row in np.asarray(m_maps):
        m = int(row[0])
        if abs(m) > Lmax:
            g[..., row[1]:row[2], :] = 0
            continue
        _, B = swsh_matrices(Ntheta, Lmax, m, s)
        ell = _ell_index(row, c.shape[-1])
        g[..., row[1]:row[2], :] = np.einsum('tk,...ik->...it', B, c[..., row[3]:row[4], ell])
    return g
