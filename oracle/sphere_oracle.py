"""Numpy restatement of the reference's S2 shallow-water IVP (BASELINE config 4).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Deliberately a DIFFERENT formulation from the product: the state is a dictionary of complex coefficient vectors per azimuthal
wavenumber m (degrees l = m .. Lmax, one entry per spin component), pencil systems are dense complex matrices solved with
numpy, transforms are plain matrix products.  The reference's packed real layout only appears in `unpack` / `pack`
(core/basis.py:2872-2914 restated) so that states can be compared with `field['c']` of the reference and of the product.

Reference code followed (paths under /root/reference/dedalus):
  transforms      core/basis.py:3062-3138 (azimuthal FFT -> spin recombination -> SWSH colatitude transform),
                  core/coords.py:219-232 (U), libraries/dedalus_sphere/sphere.py:43-64 (harmonics; via transforms_oracle)
  operators       core/basis.py:3299-3420 (divergence / gradient / Laplacian symbols), core/basis.py:3150-3152 (k),
                  core/operators.py:2125-2160 (skew: -+1j on the spin components), 2995-3050 (MulCosine: Jacobi operator Z)
  time stepping   core/timesteppers.py:552-644 (RungeKuttaIMEX.step), 95-187 (MultistepIMEX.step), tableaux via oracle/imex.py
  problem         examples/ivp_sphere_shallow_water/shallow_water.py:25-86
"""
import numpy as np
from . import imex
from . import transforms_oracle as T


def k_symbol(ell, s, mu):
    """core/basis.py:3150-3152."""
    ell = np.asarray(ell, dtype=float)
    return -mu * np.sqrt(np.maximum(0, (ell - mu * s) * (ell + mu * s + 1) / 2))


def cos_operator(Lmax, m, s):
    """cos(theta) on l = Lmin .. Lmax, Lmin = max(|m|, |s|): symmetric tridiagonal with the closed-form SWSH coefficients
        diag_l = -m s / (l (l + 1)),    off_l = sqrt((l^2 - m^2)(l^2 - s^2) / (4 l^2 - 1)) / l     (between l - 1 and l)
    (what the Jacobi operator Z of libraries/dedalus_sphere/sphere.py:91-95 evaluates to on (a, b) = (|m+s|, |m-s|))."""
    Lmin = max(abs(m), abs(s))
    ell = np.arange(Lmin, Lmax + 1, dtype=float)
    with np.errstate(divide='ignore', invalid='ignore'):
        d = np.where(ell > 0, -m * s / (ell * (ell + 1)), 0.0)
        l1 = ell[1:]
        off = np.sqrt((l1**2 - m**2) * (l1**2 - s**2) / (4 * l1**2 - 1)) / l1
    return np.diag(d) + np.diag(off, 1) + np.diag(off, -1), Lmin


class ShallowWaterOracle:
    def __init__(self, Nphi, Ntheta, dealias=1.5, R=1.0, Omega=0.0, nu=0.0, g=1.0, H=1.0):
        self.Nphi, self.Ntheta = Nphi, Ntheta
        self.Lmax = max(0, Ntheta - 2)
        self.Ngp, self.Ngt = int(np.ceil(dealias * Nphi)), int(np.ceil(dealias * Ntheta))
        self.R, self.Omega, self.nu, self.g, self.H = R, Omega, nu, g, H
        self.ms = [m for m in range(Nphi // 2) if m <= self.Lmax]
        self._mats = {}

    # ---- the reference's packed real layout <-> {m: complex (ncomp, Lmax + 1 - m)}
    def _groups(self):
        Nphi, Lmax = self.Nphi, self.Lmax
        shift = max(0, Lmax + 2 - Nphi // 2)
        shape = (Nphi // 2, Lmax + 1 + shift)
        i, j = np.indices(shape)
        m = i // 2
        ell = j - shift
        neg = ell < m
        m = np.where(neg, (Nphi // 2 - 1) - m, m)
        ell = np.where(neg, Lmax - j, ell)
        zero = i < 2
        m = np.where(zero, 0, m); ell = np.where(zero, j, ell)
        top = zero & (j > Lmax)
        m = np.where(top, Nphi // 2 - 1, m); ell = np.where(top, j - shift, ell)
        return m, ell

    def unpack(self, packed):
        """packed (ncomp, Nphi/2, Nl) real -> {m: (ncomp, Lmax + 1 - m) complex}."""
        packed = np.asarray(packed).reshape((-1,) + packed.shape[-2:])
        mg, lg = self._groups()
        out = {}
        for m in self.ms:
            c = np.zeros((packed.shape[0], self.Lmax + 1 - m), dtype=complex)
            rows, cols = np.nonzero((mg[0::2] == m) & (lg[0::2] >= m) & (lg[0::2] <= self.Lmax))
            ell = lg[2 * rows, cols]
            c[:, ell - m] = packed[:, 2 * rows, cols] + 1j * packed[:, 2 * rows + 1, cols]
            out[m] = c
        return out

    def pack(self, coeffs, ncomp):
        mg, lg = self._groups()
        packed = np.zeros((ncomp,) + mg.shape)
        for m in self.ms:
            rows, cols = np.nonzero((mg[0::2] == m) & (lg[0::2] >= m) & (lg[0::2] <= self.Lmax))
            ell = lg[2 * rows, cols]
            packed[:, 2 * rows, cols] = coeffs[m][:, ell - m].real
            packed[:, 2 * rows + 1, cols] = coeffs[m][:, ell - m].imag
        return packed

    # ---- transforms of a tensor given by its spin components
    @staticmethod
    def _U(rank):
        U1 = np.array([[-1j, 1], [1j, 1]]) / np.sqrt(2)          # coords.py:219-227, spin ordering (-, +)
        U = np.ones((1, 1))
        for _ in range(rank):
            U = np.kron(U, U1)
        return U

    def to_grid(self, coeffs, spins):
        """{m: (ncomp, nl)} spin components -> (ncomp, Ngp, Ngt) coordinate components on the dealiased grid."""
        ncomp = len(spins)
        cg = np.zeros((ncomp, self.Nphi, self.Ngt))               # (cos, -sin) interleaved along the azimuthal axis
        for m in self.ms:
            for c, s in enumerate(spins):
                _, B = T.swsh_matrices(self.Ngt, self.Lmax, m, s)
                v = B @ coeffs[m][c]
                cg[c, 2 * m], cg[c, 2 * m + 1] = v.real, v.imag
        rank = int(round(np.log2(ncomp))) if ncomp > 1 else 0
        if rank:
            z = cg[:, 0::2] + 1j * cg[:, 1::2]
            z = np.tensordot(self._U(rank).conj().T, z, axes=(1, 0))
            cg[:, 0::2], cg[:, 1::2] = z.real, z.imag
        return T.rf_backward_fft(cg, self.Ngp, axis=1)

    def from_grid(self, g, spins):
        ncomp = len(spins)
        cg = T.rf_forward_fft(g, self.Nphi, axis=1)
        rank = int(round(np.log2(ncomp))) if ncomp > 1 else 0
        z = cg[:, 0::2] + 1j * cg[:, 1::2]
        if rank:
            z = np.tensordot(self._U(rank), z, axes=(1, 0))
        out = {}
        for m in self.ms:
            c = np.zeros((ncomp, self.Lmax + 1 - m), dtype=complex)
            for q, s in enumerate(spins):
                F, _ = T.swsh_matrices(self.Ngt, self.Lmax, m, s)
                c[q] = F @ z[q, m]
            out[m] = c
        return out

    # ---- operators in coefficient space
    def _ell(self, m):
        return np.arange(m, self.Lmax + 1)

    def grad_symbol(self, m, s, mu):
        ell = self._ell(m)
        k = k_symbol(ell, s, mu) / self.R
        k[(np.abs(s) > ell) | (np.abs(s + mu) > ell)] = 0
        return k

    def lap_symbol(self, m, s):
        ell = self._ell(m)
        kl = k_symbol(ell, s + 1, -1) * k_symbol(ell, s, +1) + k_symbol(ell, s - 1, +1) * k_symbol(ell, s, -1)
        kl[np.abs(s) > ell] = 0
        return kl / self.R**2

    def matrices(self, m):
        """(M, L, valid) of  dt(u) + nu lap(lap(u)) + g grad(h) + 2 Omega MulCosine(skew(u)) ,  dt(h) + nu lap(lap(h)) + H div(u)
        for unknowns [u-, u+, h] x (l = m .. Lmax), complex."""
        if m in self._mats:
            return self._mats[m]
        nl = self.Lmax + 1 - m
        ell = self._ell(m)
        L = np.zeros((3 * nl, 3 * nl), dtype=complex)
        blk = lambda r, c: (slice(r * nl, (r + 1) * nl), slice(c * nl, (c + 1) * nl))
        for q, s in enumerate((-1, +1)):
            L[blk(q, q)] += self.nu * np.diag(self.lap_symbol(m, s) ** 2)
            L[blk(q, 2)] += self.g * np.diag(self.grad_symbol(m, 0, s))                        # grad(h): spin 0 -> s
            C, Lmin = cos_operator(self.Lmax, m, s)
            full = np.zeros((nl, nl)); full[Lmin - m:, Lmin - m:] = C
            L[blk(q, q)] += 2 * self.Omega * full * (1j * s)                                   # MulCosine(skew(u))
            L[blk(2, q)] += self.H * np.diag(self.grad_symbol(m, s, -s))                       # div(u): spin s -> 0
        L[blk(2, 2)] += self.nu * np.diag(self.lap_symbol(m, 0) ** 2)
        valid = np.concatenate([ell >= max(m, 1), ell >= max(m, 1), ell >= m])
        M = np.diag(valid.astype(complex))
        L = L * np.outer(valid, valid) + np.diag((~valid).astype(complex))
        self._mats[m] = (M, L, valid)
        return self._mats[m]

    def rhs(self, X):
        """F = [-(u . grad u), -div(h u)] evaluated through the dealiased grid (shallow_water.py:82-83)."""
        u = {m: X[m][:2] for m in self.ms}
        h = {m: X[m][2:3] for m in self.ms}
        gradu = {}
        for m in self.ms:
            rows = []
            for mu in (-1, +1):
                for q, s in enumerate((-1, +1)):
                    rows.append(self.grad_symbol(m, s, mu) * u[m][q])
            gradu[m] = np.array(rows)
        ug = self.to_grid(u, (-1, +1))
        hg = self.to_grid(h, (0,))
        gg = self.to_grid(gradu, (-2, 0, 0, 2)).reshape(2, 2, self.Ngp, self.Ngt)
        adv = np.einsum('i...,ij...->j...', ug, gg)                 # u @ grad(u): contracts u with the FIRST index of grad(u)
        flux = hg * ug
        Fu = self.from_grid(adv, (-1, +1))
        Fh_vec = self.from_grid(flux, (-1, +1))
        out = {}
        for m in self.ms:
            div = sum(self.grad_symbol(m, s, -s) * Fh_vec[m][q] for q, s in enumerate((-1, +1)))
            out[m] = np.concatenate([-Fu[m], -div[None, :]], axis=0)
        return out

    # ---- IMEX loops
    def run(self, u_packed, h_packed, steps, dt, scheme="RK222"):
        uc, hc = self.unpack(u_packed), self.unpack(h_packed)
        X = {m: np.concatenate([uc[m], hc[m]], axis=0) for m in self.ms}
        flat = lambda V, m: V[m].reshape(-1)
        dt_seq = [float(dt)] * steps if np.isscalar(dt) else [float(v) for v in dt]
        if scheme in imex.RK:
            tab = imex.RK[scheme]
            A, Hh, stages = tab['A'], tab['H'], len(tab['c']) - 1
            for dt in dt_seq:
                MX0 = {m: self.matrices(m)[0] @ flat(X, m) for m in self.ms}
                LX, F = [], []
                for i in range(1, stages + 1):
                    LX.append({m: self.matrices(m)[1] @ flat(X, m) for m in self.ms})
                    Fi = self.rhs(X)
                    F.append({m: flat(Fi, m) * self.matrices(m)[2] for m in self.ms})
                    for m in self.ms:
                        Mm, Lm, _ = self.matrices(m)
                        rhs = MX0[m].copy()
                        for j in range(i):
                            rhs += dt * A[i, j] * F[j][m] - dt * Hh[i, j] * LX[j][m]
                        X[m] = np.linalg.solve(Mm + dt * Hh[i, i] * Lm, rhs).reshape(3, -1)
        elif scheme == "SBDF2":
            MXh, LXh, Fh = [], [], []
            prev = dt_seq[0]
            for it, dt in enumerate(dt_seq):
                a, b, c = imex.sbdf2(dt, prev, it)
                prev = dt
                MXh.insert(0, {m: self.matrices(m)[0] @ flat(X, m) for m in self.ms})
                LXh.insert(0, {m: self.matrices(m)[1] @ flat(X, m) for m in self.ms})
                Fi = self.rhs(X)
                Fh.insert(0, {m: flat(Fi, m) * self.matrices(m)[2] for m in self.ms})
                for m in self.ms:
                    Mm, Lm, _ = self.matrices(m)
                    rhs = np.zeros(Mm.shape[0], dtype=complex)
                    for j in range(1, len(c)):
                        if c[j] != 0 and j - 1 < len(Fh):
                            rhs += c[j] * Fh[j - 1][m]
                    for j in range(1, len(a)):
                        if a[j] != 0 and j - 1 < len(MXh):
                            rhs -= a[j] * MXh[j - 1][m]
                    for j in range(1, len(b)):
                        if b[j] != 0 and j - 1 < len(LXh):
                            rhs -= b[j] * LXh[j - 1][m]
                    X[m] = np.linalg.solve(a[0] * Mm + b[0] * Lm, rhs).reshape(3, -1)
                del MXh[2:], LXh[2:], Fh[2:]
        else:
            raise ValueError(scheme)
        u = self.pack({m: X[m][:2] for m in self.ms}, 2)
        h = self.pack({m: X[m][2:3] for m in self.ms}, 1)[0]
        return dict(u=u, h=h)


def shallow_water_parameters():
    """Units and constants of examples/ivp_sphere_shallow_water/shallow_water.py:25-38."""
    meter = 1 / 6.37122e6
    second = 1 / 3600
    return dict(R=6.37122e6 * meter, Omega=7.292e-5 / second, nu=1e5 * meter**2 / second / 32**2,
                g=9.80616 * meter / second**2, H=1e4 * meter)


def run(Nphi, Ntheta, u0_c, h0_c, steps, dt, scheme="RK222", dealias=1.5):
    orc = ShallowWaterOracle(Nphi, Ntheta, dealias=dealias, **shallow_water_parameters())
    return orc.run(u0_c, h0_c, steps, dt, scheme)
