"""Spherical-shell basis: the transform chain of scalar / vector / tensor fields.

Reference: core/basis.py:4336-4508 (ShellBasis; forward / backward_transform_radius 4474-4508), 3684-3850 (ShellRadialBasis),
3422-3627 (RegularityBasis: regularity classes and recombination), core/coords.py:313-385 (SphericalCoordinates),
libraries/dedalus_sphere/spin_operators.py:276-358 (the spin <-> regularity intertwiner Q(l)).

Layouts (real dtype): grid (comps, Nphi_g, Ntheta_g, Nr_g), coordinate components (phi, theta, r); coefficients
(comps, Nphi/2, Lmax + 1 + shift, Nr), regularity components (-, +, 0), the angular part in the sphere's folded triangular (m, l)
packing (dedalus_b200/sphere.py).  Chain towards the grid, every stage one launch over all components:
  radial Jacobi transform along the contiguous last axis (Chebyshev-grid fast transform, csrc/rfft_regs.cu k_chbwd_regs)
  -> regularity -> spin recombination Q(l) (csrc/banded.cu k_pair_lincomb with a per-(m, l) symbol table, constant along r)
  -> SWSH colatitude transform per spin weight (csrc/pointwise.cu k_ragged_matvec, the radial index as the trailing batch axis)
  -> spin -> component recombination (k_pair_lincomb) -> azimuthal real FFT.
Only the field transforms are built: the shell's operators and per-l pencil systems (config 5) are not.
"""
import itertools
import numpy as np
from .basis import Basis
from .coords import SphericalCoordinates
from .sphere import SphereBasis, PairProgram


class Intertwiner:
    """Q(l)[spin, regularity]: orthogonal map from regularity to spin components of rank-n tensors (indexing -, +, 0).
    Restates the recursion of libraries/dedalus_sphere/spin_operators.py:310-358."""

    indexing = (-1, +1, 0)

    def __init__(self, L):
        self.L = L
        self._memo = {}

    def k(self, mu, s):
        return -mu * np.sqrt((self.L - s * mu) * (self.L + s * mu + 1) / 2)

    def forbidden_spin(self, spin):
        return self.L < abs(sum(spin))

    def forbidden_regularity(self, reg):
        if self.L >= len(reg):
            return False
        walk = (self.L,)
        for r in reg[::-1]:
            walk += (walk[-1] + r,)
            if walk[-1] < 0 or walk[-2:] == (0, 0):
                return True
        return False

    def element(self, spin, reg):
        key = (spin, reg)
        if key in self._memo:
            return self._memo[key]
        if len(spin) == 0:
            return 1.0
        if self.forbidden_spin(spin) or self.forbidden_regularity(reg):
            self._memo[key] = 0.0
            return 0.0
        sigma, a = spin[0], reg[0]
        tau, b = spin[1:], reg[1:]
        R = 0.0
        for i, t in enumerate(tau):
            if t + sigma == 0:
                R -= self.element(tau[:i] + (0,) + tau[i + 1:], b)
            if t == 0:
                R += self.element(tau[:i] + (sigma,) + tau[i + 1:], b)
        Q = self.element(tau, b)
        R -= self.k(sigma, sum(tau)) * Q
        J = self.L + sum(b)
        if sigma != 0:
            Q = 0.0
        if a == -1:
            val = (Q * J - R) / np.sqrt(J * (2 * J + 1))
        elif a == 0:
            val = sigma * R / np.sqrt(J * (J + 1))
        else:
            val = (Q * (J + 1) + R) / np.sqrt((J + 1) * (2 * J + 1))
        self._memo[key] = float(val)
        return self._memo[key]

    def matrix(self, rank):
        idx = list(itertools.product(self.indexing, repeat=rank))
        return np.array([[self.element(s, r) for r in idx] for s in idx])


class ShellBasis(Basis):
    dim = 3
    kind = "Shell"

    def __init__(self, coordsys, shape, dtype=np.float64, radii=(1, 2), k=0, alpha=(-0.5, -0.5), dealias=(1, 1, 1),
                 azimuth_library=None, colatitude_library=None, radius_library=None):
        if not isinstance(coordsys, SphericalCoordinates):
            raise ValueError("Shell coordsys must be SphericalCoordinates.")
        shape = tuple(int(n) for n in shape)
        if len(shape) != 3:
            raise ValueError("Shell shape must have length 3.")
        radii = tuple(radii)
        if len(radii) != 2:
            raise ValueError("Shell radii must have length 2")
        if min(radii) <= 0:
            raise ValueError("Shell radii must be positive.")
        if radii[0] >= radii[1]:
            raise ValueError("Shell radii must be in increasing order.")
        if isinstance(alpha, (int, float)):
            alpha = (alpha,) * 2
        alpha = tuple(alpha)
        if isinstance(dealias, (int, float)):
            dealias = (dealias,) * 3
        dealias = tuple(dealias)
        if len(dealias) != 3:
            raise ValueError("Shell dealias must have length 3.")
        if np.dtype(dtype) != np.float64:
            raise NotImplementedError("only real (float64) shell fields are built")
        if alpha != (-0.5, -0.5):
            raise NotImplementedError("only the Chebyshev radial grid (alpha = -1/2) is built")
        self.coordsys, self.coord = coordsys, coordsys.coords[0]
        self.shape, self.dtype, self.radii, self.k, self.alpha, self.dealias = shape, np.float64, radii, int(k), alpha, dealias
        self.volume = 4 / 3 * np.pi * (radii[1]**3 - radii[0]**3)
        self.dR = radii[1] - radii[0]
        self.rho = (radii[1] + radii[0]) / self.dR
        self.sphere_basis = SphereBasis(coordsys.S2coordsys, shape[:2], radius=1, dealias=dealias[:2])
        self.Lmax = self.sphere_basis.Lmax
        self.coeff_shape = self.sphere_basis.coeff_shape + (shape[2],)
        self._key = (coordsys, shape, radii, self.k, alpha, dealias)
        self._ctor_args = dict(coordsys=coordsys, shape=shape, dtype=np.float64, radii=radii, k=self.k, alpha=alpha, dealias=dealias)
        self._plans = {}
        self.grid_params = (coordsys, radii, alpha, dealias)

    # ---- related bases (reference basis.py:4401-4413, 4236-4245)
    @property
    def radial_basis(self):
        if 'radial' not in self._plans:
            self._plans['radial'] = ShellRadialBasis(self.coordsys, self.shape[2], radii=self.radii, alpha=self.alpha,
                                                     dealias=(self.dealias[2],), k=self.k)
        return self._plans['radial']

    def S2_basis(self, radius=None):
        radius = max(self.radii) if radius is None else radius
        key = ('S2', radius)
        if key not in self._plans:
            self._plans[key] = SphereBasis(self.coordsys, self.shape[:2], radius=radius, dealias=self.dealias[:2])
        return self._plans[key]

    @property
    def outer_surface(self):
        return self.S2_basis(self.radii[1])

    @property
    def inner_surface(self):
        return self.S2_basis(self.radii[0])

    @classmethod
    def _make(cls, **kw):
        return cls(**kw)

    def __repr__(self):
        return f"ShellBasis({self.shape}, radii={self.radii}, k={self.k})"

    def axis_size(self, sub=0):
        return self.coeff_shape[sub]

    def axis_grid_size(self, scale, sub=0):
        return int(np.ceil(scale * self.shape[sub]))

    def axis_dealias(self, sub=0):
        return self.dealias[sub]

    def axis_group_size(self, sub=0):
        return (2, 1, 1)[sub]

    def grid_shape(self, scales):
        return tuple(self.axis_grid_size(s, i) for i, s in enumerate(scales))

    # ---- basis algebra (reference basis.py:4423-4462, 3718-3762): sums take the larger k, products add them
    def __add__(self, other):
        if other is None or other == self:
            return self
        if isinstance(other, ShellBasis) and self.grid_params == other.grid_params and self.shape == other.shape:
            return self.clone_with(k=max(self.k, other.k))
        if isinstance(other, ShellRadialBasis) and self.grid_params[:3] == other.grid_params[:3]:
            return self.clone_with(k=max(self.k, other.k))
        if isinstance(other, SphereBasis):
            return self
        return NotImplemented
    __radd__ = __add__

    def __mul__(self, other):
        if other is None:
            return self
        if isinstance(other, (ShellBasis, ShellRadialBasis)) and self.grid_params[:3] == other.grid_params[:3]:
            return self.clone_with(k=self.k + other.k)
        if isinstance(other, SphereBasis):
            return self
        return NotImplemented
    __rmul__ = __mul__

    def derivative_basis(self, order=1):
        return self.clone_with(k=self.k + order)

    # ---- grids (reference basis.py:3769-3774, 4290-4300)
    def global_grid_radius(self, scale):
        from . import jacobi
        z = jacobi.gauss_grid(self.axis_grid_size(scale, 2), self.alpha[0], self.alpha[1])[0]
        return self.dR / 2 * (z + self.rho)

    def local_grids(self, dist, scales):
        ax = dist.get_basis_axis(self)
        sb = self.sphere_basis
        out = []
        for sub, g in enumerate((sb.global_grid_azimuth(scales[0]), sb.global_grid_colatitude(scales[1]), self.global_grid_radius(scales[2]))):
            g = g[dist.grid_local_slice(ax + sub, self, scales[sub])]
            shp = [1] * dist.dim
            shp[ax + sub] = g.size
            out.append(g.reshape(shp))
        return tuple(out)

    # ---- tensor bookkeeping
    def spin_weights(self, tensorsig):
        S = np.zeros(tuple(cs.dim for cs in tensorsig), dtype=int)
        for i, cs in enumerate(tensorsig):
            if cs is not self.coordsys:
                raise NotImplementedError("tensor indices over other coordinate systems on a shell basis")
            shp = [1] * len(tensorsig); shp[i] = 3
            S = S + np.array(cs.spin_ordering).reshape(shp)
        return S

    def radial_plan(self, Nr_g):
        key = ('rad', int(Nr_g))
        if key not in self._plans:
            from .transforms import FastChebyshevTransform
            a, b = self.alpha[0] + self.k, self.alpha[1] + self.k
            self._plans[key] = FastChebyshevTransform(Nr_g, self.shape[2], a, b, self.alpha[0], self.alpha[1])
        return self._plans[key]

    def spin_table(self, rank, forward, device):
        key = ('spin', rank, bool(forward), str(device))
        if key not in self._plans:
            U = SphericalCoordinates.U_forward(rank) if forward else SphericalCoordinates.U_backward(rank)
            self._plans[key] = PairProgram.from_matrix(U, device)
        return self._plans[key]

    def regularity_table(self, rank, forward, device, dist=None):
        """db_pair_lincomb program of the regularity recombination: out = Q(l)^T in (forward, spin -> regularity) or Q(l) in
        (backward), one real symbol table over the (m, l) packing per nonzero (out, in) pair (reference basis.py:3590-3627)."""
        j0, j1 = self.sphere_basis.local_pairs(dist)
        key = ('reg', rank, bool(forward), str(device), j0, j1)
        if key not in self._plans:
            import torch
            sb = self.sphere_basis
            _, ell_map = sb.elements_to_groups()
            ell_pairs = ell_map[0::2][j0:j1]                            # (local pairs, Nl)
            n = 3 ** rank
            Q = np.zeros((sb.Lmax + 1, n, n))
            for ell in range(sb.Lmax + 1):
                Q[ell] = Intertwiner(ell).matrix(rank)
            if forward:
                Q = Q.transpose(0, 2, 1)
            in_range = ell_pairs <= sb.Lmax
            lidx = np.minimum(ell_pairs, sb.Lmax)
            syms, rows = [], []
            for o in range(n):
                row = []
                for i in range(n):
                    if np.any(Q[:, o, i] != 0):
                        row.append((i, 1.0, len(syms) * ell_pairs.size))
                        syms.append(np.where(in_range, Q[lidx, o, i], 0.0).ravel())
                rows.append(row)
            self._plans[key] = PairProgram(rows, device, torch.from_numpy(np.concatenate(syms)).to(device))
        return self._plans[key]


class ShellRadialBasis(Basis):
    """Radial part of a shell basis as a 1-D basis on the radius coordinate: (dR / r)^k P_n^(alpha + k)(z) (reference
    basis.py:3684-3850).  Fields on it (er, rvec, ...) are the radial non-constant coefficients of shell problems; their
    transforms (regularity components at l = 0) are tiny host-side setup work (dedalus_b200/shell_ivp.py)."""
    dim = 1
    kind = "ShellRadial"
    group_size = 1

    def __init__(self, coordsys, radial_size, dtype=np.float64, radii=(1, 2), alpha=(-0.5, -0.5), dealias=(1,), k=0, radius_library=None):
        self.coordsys, self.coord = coordsys, coordsys.radius
        self.size = int(radial_size)
        self.radii, self.alpha, self.k = tuple(radii), tuple(alpha), int(k)
        self.dealias = (dealias,) if isinstance(dealias, (int, float)) else tuple(dealias)
        self.dtype = np.float64
        self.dR = radii[1] - radii[0]
        self.rho = (radii[1] + radii[0]) / self.dR
        self.grid_params = (coordsys, self.radii, self.alpha)
        self._key = (coordsys, self.size, self.radii, self.alpha, self.dealias, self.k)
        self._ctor_args = dict(coordsys=coordsys, radial_size=self.size, radii=self.radii, alpha=self.alpha, dealias=self.dealias, k=self.k)

    @classmethod
    def _make(cls, **kw):
        return cls(**kw)

    def grid_size(self, scale):
        return int(np.ceil(scale * self.size))

    def _native_grid(self, scale):
        from . import jacobi
        return jacobi.gauss_grid(self.grid_size(scale), self.alpha[0], self.alpha[1])[0]

    def global_grid(self, scale=None):
        scale = self.dealias[0] if scale is None else scale
        return self.dR / 2 * (self._native_grid(scale) + self.rho)

    def derivative_basis(self, order=1):
        return self.clone_with(k=self.k + order)

    def __add__(self, other):
        if other is None or other == self:
            return self
        if isinstance(other, ShellRadialBasis) and self.grid_params == other.grid_params:
            return self.clone_with(k=max(self.k, other.k))
        if isinstance(other, ShellBasis):
            return other + self
        return NotImplemented
    __radd__ = __add__

    def __mul__(self, other):
        if other is None:
            return self
        if isinstance(other, ShellRadialBasis) and self.grid_params == other.grid_params:
            return self.clone_with(k=self.k + other.k)
        if isinstance(other, ShellBasis):
            return other * self
        return NotImplemented
    __rmul__ = __mul__


def shell_basis_of(field_or_bases):
    """The shell basis of a field or expression.  The per-axis bases of a product with a radial-basis coefficient differ in k along the
    radial axis only (k adds there); the entry of the radial axis -- the largest k -- is the basis of the whole."""
    bases = getattr(field_or_bases, 'bases', field_or_bases)
    found = [b for b in bases if isinstance(b, ShellBasis)]
    return max(found, key=lambda b: b.k) if found else None


def _spin_groups(spins):
    c = 0
    while c < len(spins):
        c1 = c
        while c1 < len(spins) and spins[c1] == spins[c]:
            c1 += 1
        yield spins[c], c, c1
        c = c1


def shell_components_to_grid(basis, c, rank, scales, dist=None):
    """Regularity components (ncomp, local Nphi/2 rows, Nl, Nr) in basis `basis` (its k) -> coordinate components on the
    (colatitude-distributed) grid."""
    import torch
    sb = basis.sphere_basis
    ncomp, Nc0 = c.shape[0], c.shape[1]
    spins = ([int(s) for s in basis.spin_weights((basis.coordsys,) * rank).reshape(-1)] if rank else [0])
    Ngp, Ngt, Ngr = basis.grid_shape(scales)
    Nc1, Nr = basis.coeff_shape[1:]
    rows = 2 * len(sb.local_wavenumbers(dist))
    dev, dt = c.device, c.dtype
    cr = torch.empty((ncomp, Nc0, Nc1, Ngr), dtype=dt, device=dev)
    basis.radial_plan(Ngr).backward(c, cr, 3)
    if rank > 0:
        tmp = torch.empty_like(cr)
        basis.regularity_table(rank, False, dev, dist).apply(cr, tmp, Nc0 // 2, Nc1 * Ngr, sym_div=Ngr)
        cr = tmp
    if basis.k > 0:
        cr = cr * torch.from_numpy((basis.dR / basis.global_grid_radius(scales[2])) ** basis.k).to(dev)
    cg = torch.empty((ncomp, rows, Ngt, Ngr), dtype=dt, device=dev)
    for s, c0, c1 in _spin_groups(spins):
        sb.colatitude_plan(Ngt, s, dist).backward(cr[c0:c1], cg[c0:c1], 2)
    if rank > 0:
        tmp = torch.empty_like(cg)
        basis.spin_table(rank, False, dev).apply(cg, tmp, rows // 2, Ngt * Ngr)
        cg = tmp
    if dist is not None and dist.size > 1:
        cg = sb.hop(dist).to_grid_side(cg).contiguous()
    g = torch.empty((ncomp, Ngp, cg.shape[2], Ngr), dtype=dt, device=dev)
    sb.azimuth_plan(Ngp).backward(cg, g, 1)
    return g


def shell_grid_to_components(basis, g, rank, dist=None):
    """Coordinate components on the grid (ncomp, Ngp, local Ngt, Ngr) -> regularity components in basis `basis`."""
    import torch
    sb = basis.sphere_basis
    ncomp, Ngp, tb, Ngr = g.shape
    spins = ([int(s) for s in basis.spin_weights((basis.coordsys,) * rank).reshape(-1)] if rank else [0])
    Nc1, Nr = basis.coeff_shape[1:]
    j0, j1 = sb.local_pairs(dist)
    Nc0 = 2 * (j1 - j0)
    Nphi = basis.shape[0]
    dev, dt = g.device, g.dtype
    cg = torch.empty((ncomp, Nphi, tb, Ngr), dtype=dt, device=dev)
    sb.azimuth_plan(Ngp).forward(g, cg, 1)
    if dist is not None and dist.size > 1:
        cg = sb.hop(dist).to_coeff_side(cg).contiguous()
    rows, Ngt = cg.shape[1], cg.shape[2]
    if rank > 0:
        tmp = torch.empty_like(cg)
        basis.spin_table(rank, True, dev).apply(cg, tmp, rows // 2, Ngt * Ngr)
        cg = tmp
    cr = torch.zeros((ncomp, Nc0, Nc1, Ngr), dtype=dt, device=dev)
    for s, c0, c1 in _spin_groups(spins):
        sb.colatitude_plan(Ngt, s, dist).forward(cg[c0:c1], cr[c0:c1], 2)
    if basis.k > 0:
        scale = Ngr / basis.shape[2]
        cr = cr * torch.from_numpy((basis.dR / basis.global_grid_radius(scale)) ** (-basis.k)).to(dev)
    if rank > 0:
        tmp = torch.empty_like(cr)
        basis.regularity_table(rank, True, dev, dist).apply(cr, tmp, Nc0 // 2, Nc1 * Ngr, sym_div=Ngr)
        cr = tmp
    c = torch.empty((ncomp, Nc0, Nc1, Nr), dtype=dt, device=dev)
    basis.radial_plan(Ngr).forward(cr.contiguous(), c, 3)
    return c


def transform_shell_field(field, layout):
    """field['c'] <-> field['g'] for fields on a ShellBasis (single GPU; reference basis.py:4474-4508 + the sphere chain)."""
    basis = shell_basis_of(field)
    if any(b is not None and b is not basis for b in field.bases):
        raise NotImplementedError("shell basis combined with other bases")
    ax = field.dist.get_basis_axis(basis)
    scales = field.scales[ax:ax + 3]
    rank = len(field.tensorsig)
    ncomp = max(field.ncomp, 1)
    data = field.device_data()
    nt = len(field.tshape)
    if layout == 'g':
        g = shell_components_to_grid(basis, data.reshape((ncomp,) + tuple(data.shape[nt:])).contiguous(), rank, scales, field.dist)
        field.set_device_data(g.reshape(field.tshape + tuple(g.shape[1:])), 'g')
    else:
        c = shell_grid_to_components(basis, data.reshape((ncomp,) + tuple(data.shape[nt:])).contiguous(), rank, field.dist)
        field.set_device_data(c.reshape(field.tshape + tuple(c.shape[1:])), 'c')


def transform_radial_field(field, layout):
    """Fields on a ShellRadialBasis alone (er, rvec: the radial coefficients of shell problems): coordinate components on the
    radial grid <-> regularity components (at l = 0) in Jacobi coefficients.  Host-side numpy: these are a few 1-D arrays used
    while the pencil matrices are assembled, never inside a time step (reference: RegularityBasis transforms with the azimuthal
    and colatitude axes constant, core/basis.py:3629-3660, 3814-3846)."""
    from . import jacobi
    basis = field.bases[-1]
    rank = len(field.tensorsig)
    ncomp = max(field.ncomp, 1)
    data = np.asarray(field.data).reshape(ncomp, -1)
    a, b = basis.alpha[0] + basis.k, basis.alpha[1] + basis.k
    N = basis.grid_size(field.scales[-1])
    z, w = jacobi.gauss_grid(N, basis.alpha[0], basis.alpha[1])
    r = basis.dR / 2 * (z + basis.rho)
    if rank:
        cs = field.tensorsig[0]
        U, Q = cs.U_forward(rank), Intertwiner(0).matrix(rank)
    if layout == 'g':
        P = jacobi.polynomials(basis.size, a, b, z)                         # (Nr, N)
        reg = (data @ P) * (basis.dR / r) ** basis.k
        out = reg if not rank else (U.conj().T @ (Q @ reg)).real
        shape = field.tshape + (1,) * (field.dist.dim - 1) + (N,)
    else:
        spin = data if not rank else (U @ data)
        if rank and np.abs(spin.imag).max() > 1e-14 * max(np.abs(spin).max(), 1e-300):
            raise NotImplementedError("radial fields with angular components")
        reg = np.real(spin) if not rank else Q.T @ np.real(spin)
        reg = reg * (basis.dR / r) ** (-basis.k)
        # grid (alpha0 weight) -> coefficients of the (alpha + k) polynomials: project in the alpha0 basis, then convert
        P0 = jacobi.polynomials(max(basis.size, N), basis.alpha[0], basis.alpha[1], z)[:, :] * w[None, :]
        c0 = reg @ P0.T
        c0[:, N:] = 0
        C = jacobi.conversion_matrix(c0.shape[1], basis.alpha[0], basis.alpha[1], a, b)
        out = (C @ c0.T).T[:, :basis.size]
        shape = field.tshape + (1,) * (field.dist.dim - 1) + (basis.size,)
    field.layout = layout
    field._host = np.ascontiguousarray(out.reshape(shape))
    field._dev = None
    field._fresh = 'host'
