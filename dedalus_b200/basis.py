"""Spectral bases for the Cartesian hot path: RealFourier, ComplexFourier, Jacobi (ChebyshevT/U/V, Legendre).

Host-side description only: sizes, grids, wavenumbers and the 1-D operator matrices used for pencil
assembly.  Mirrors the public constructor signatures and conventions of the reference
(dedalus/core/basis.py:435-660 Jacobi, 817-936 FourierBase, 1108-1183 RealFourier, 942-1100 ComplexFourier).
Transforms themselves live in dedalus_b200/transforms.py (CUDA).
"""
import numpy as np
from scipy import sparse
from . import jacobi
from .coords import Coordinate, AffineCOV


class Basis:
    dim = 1

    # per-axis accessors (sub = axis index inside a multi-dimensional basis, always 0 here; dedalus_b200/sphere.py overrides)
    def axis_size(self, sub=0):
        return self.size

    def axis_grid_size(self, scale, sub=0):
        return self.grid_size(scale)

    def axis_dealias(self, sub=0):
        return self.dealias[0]

    def axis_group_size(self, sub=0):
        return self.group_size

    def grid_size(self, scale):
        """ceil(scale * size), as the reference rounds (core/basis.py:141-144); size-1 bases stay at 1."""
        if self.size == 1:
            return 1
        g = float(scale) * self.size
        return int(np.ceil(g - 1e-12))

    def global_grid(self, dist_or_scale=None, scale=None):
        """global_grid(scale): the 1-D grid; global_grid(dist, scale): the reference's form (core/basis.py:364-368), shaped for
        broadcasting along this basis' axis of the distributor."""
        dist = None
        if dist_or_scale is not None and not isinstance(dist_or_scale, (int, float, np.integer, np.floating)):
            dist = dist_or_scale
        elif scale is None:
            scale = dist_or_scale
        scale = self.dealias[0] if scale is None else scale
        grid = self.COV.problem_coord(self._native_grid(scale))
        if dist is None:
            return grid
        shape = [1] * dist.dim
        shape[dist.get_basis_axis(self)] = grid.size
        return grid.reshape(shape)

    def local_grid(self, dist, scale=None):
        """The reference's basis.local_grid(dist, scale) (core/basis.py:374-380)."""
        return dist.local_grid(self, scale)

    def clone_with(self, **kw):
        args = dict(self._ctor_args)
        args.update(kw)
        return type(self)._make(**args)

    def __eq__(self, other):
        return type(self) is type(other) and self._key == other._key

    def __hash__(self):
        return hash((type(self).__name__, self._key))


class FourierBase(Basis):
    native_bounds = (0, 2 * np.pi)

    def __init__(self, coord, size, bounds, dealias=1, library=None):
        if not isinstance(coord, Coordinate):
            raise ValueError("Fourier coord must be Coordinate object.")
        size = int(size)
        if size <= 0:
            raise ValueError("Fourier size must be positive.")
        bounds = tuple(bounds)
        if len(bounds) != 2:
            raise ValueError("Fourier bounds must have length 2.")
        dealias = (dealias,) if isinstance(dealias, (int, float)) else tuple(dealias)
        self.coord, self.size, self.bounds, self.dealias, self.library = coord, size, bounds, dealias, library
        self.COV = AffineCOV(self.native_bounds, bounds)
        self.constant_mode_value = 1
        self._ctor_args = dict(coord=coord, size=size, bounds=bounds, dealias=dealias, library=library)
        self._key = (coord, size, bounds, dealias)

    @classmethod
    def _make(cls, **kw):
        return cls(**kw)

    def _native_grid(self, scale):
        N = self.grid_size(scale)
        return (2 * np.pi / N) * np.arange(N)

    @property
    def wavenumbers(self):
        return self.native_wavenumbers / self.COV.stretch

    # basis algebra (reference basis.py:866-885)
    def __add__(self, other):
        if other is None or other == self:
            return self
        return NotImplemented
    __radd__ = __add__
    __mul__ = __add__
    __rmul__ = __add__

    def derivative_basis(self, order=1):
        return self


class RealFourier(FourierBase):
    """cos / -sin modes interleaved: [cos 0x, -sin 0x, cos 1x, -sin 1x, ...] (reference basis.py:1108-1134)."""
    group_size = 2
    kind = "RealFourier"

    @property
    def native_wavenumbers(self):
        kmax = (self.size - 1) // 2
        return np.repeat(np.arange(0, kmax + 1), 2)

    @property
    def n_groups(self):
        return self.size // 2

    def group_wavenumber(self, g):
        return g / self.COV.stretch

    # --- separable-axis symbols: dict {power m: small matrix}, meaning sum_m k^m Mat_m (k = physical wavenumber)
    def sym_identity(self):
        return {0: np.eye(2)}

    def sym_derivative(self):
        # d/dx cos(kx) = k * (-sin kx);  d/dx (-sin kx) = -k cos(kx)   (reference basis.py:1217-1224)
        return {1: np.array([[0., -1.], [1., 0.]])}

    def sym_embed_constant(self):
        return {0: np.array([[1.], [0.]])}

    def sym_integrate(self):
        return {0: np.array([[self.COV.problem_length, 0.]])}

    def sym_average(self):
        return {0: np.array([[1., 0.]])}

    # --- full matrices (used when this is the coupled last axis)
    def derivative_matrix(self):
        k = self.wavenumbers[::2]
        blocks = [sparse.coo_matrix(np.array([[0., -kk], [kk, 0.]])) for kk in k]
        return sparse.block_diag(blocks, format='csr')

    def embed_constant_vector(self):
        v = np.zeros((self.size, 1)); v[0, 0] = 1
        return sparse.csr_matrix(v)

    def integration_vector(self):
        v = np.zeros((1, self.size)); v[0, 0] = self.COV.problem_length
        return sparse.csr_matrix(v)

    def interpolation_vector(self, position):
        x = self.COV.native_coord(position)
        k = self.native_wavenumbers
        v = np.zeros(k.size)
        v[0::2] = np.cos(k[0::2] * x)
        v[1::2] = -np.sin(k[1::2] * x)
        return sparse.csr_matrix(v[None, :])

    def valid_coeff_mask(self):
        m = np.ones(self.size, dtype=bool)
        m[1] = False      # -sin(0 x)
        return m


class ComplexFourier(FourierBase):
    """exp(i k x) modes ordered [0..kmax, (Nyquist), -kmax..-1] (reference basis.py:942-960)."""
    group_size = 1
    kind = "ComplexFourier"

    @property
    def native_wavenumbers(self):
        kmax = (self.size - 1) // 2
        k = np.concatenate((np.arange(0, kmax + 1), np.arange(-kmax, 0)))
        if self.size % 2 == 0:
            k = np.insert(k, kmax + 1, kmax + 1)  # placeholder slot for the dropped Nyquist mode
        return k

    def derivative_matrix(self):
        return sparse.diags([1j * self.wavenumbers], [0], format='csr')

    def embed_constant_vector(self):
        v = np.zeros((self.size, 1), dtype=complex); v[0, 0] = 1
        return sparse.csr_matrix(v)

    def integration_vector(self):
        v = np.zeros((1, self.size), dtype=complex); v[0, 0] = self.COV.problem_length
        return sparse.csr_matrix(v)

    def interpolation_vector(self, position):
        x = self.COV.native_coord(position)
        return sparse.csr_matrix(np.exp(1j * self.native_wavenumbers * x)[None, :])

    def valid_coeff_mask(self):
        m = np.ones(self.size, dtype=bool)
        if self.size % 2 == 0:
            m[(self.size - 1) // 2 + 1] = False   # Nyquist
        return m

    def sym_identity(self):
        return {0: np.eye(1)}

    def sym_derivative(self):
        return {1: np.array([[1j]])}


class Jacobi(Basis):
    """Jacobi polynomial basis, unit-weight normalisation (reference basis.py:435-633)."""
    native_bounds = (-1, 1)
    group_size = 1
    kind = "Jacobi"

    def __init__(self, coord, size, bounds, a, b, a0=None, b0=None, dealias=1, library=None):
        if not isinstance(coord, Coordinate):
            raise ValueError("Jacobi coord must be Coordinate object.")
        size = int(size)
        if size <= 0:
            raise ValueError("Jacobi size must be positive.")
        bounds = tuple(bounds)
        if len(bounds) != 2:
            raise ValueError("Jacobi bounds must have length 2.")
        a, b = float(a), float(b)
        a0 = a if a0 is None else float(a0)
        b0 = b if b0 is None else float(b0)
        dealias = (dealias,) if isinstance(dealias, (int, float)) else tuple(dealias)
        if library is None:
            library = "b200_dct" if (a0 == b0 == -0.5) else "b200_matrix"
        self.coord, self.size, self.bounds = coord, size, bounds
        self.a, self.b, self.a0, self.b0, self.dealias, self.library = a, b, a0, b0, dealias, library
        self.COV = AffineCOV(self.native_bounds, bounds)
        self.constant_mode_value = 1 / np.sqrt(jacobi.mass(a, b))
        self._ctor_args = dict(coord=coord, size=size, bounds=bounds, a=a, b=b, a0=a0, b0=b0, dealias=dealias, library=library)
        self._key = (coord, size, bounds, a, b, a0, b0, dealias)
        self.grid_params = (coord, bounds, a0, b0, dealias)

    @classmethod
    def _make(cls, **kw):
        return cls(**kw)

    def __repr__(self):
        return f"Jacobi({self.coord.name},{self.size},a0={self.a0},b0={self.b0},a={self.a},b={self.b})"

    def _native_grid(self, scale):
        return jacobi.gauss_grid(self.grid_size(scale), self.a0, self.b0)[0]

    # basis algebra (reference basis.py:522-569)
    def __add__(self, other):
        if other is None or other == self:
            return self
        if isinstance(other, Jacobi) and self.grid_params == other.grid_params:
            return self.clone_with(size=max(self.size, other.size), a=max(self.a, other.a), b=max(self.b, other.b))
        return NotImplemented
    __radd__ = __add__

    def __mul__(self, other):
        if other is None or other == self:
            return self
        if isinstance(other, Jacobi) and self.grid_params == other.grid_params:
            return self.clone_with(size=max(self.size, other.size), a=self.a0, b=self.b0)
        return NotImplemented
    __rmul__ = __mul__

    def derivative_basis(self, order=1):
        return self.clone_with(a=self.a + order, b=self.b + order)

    # full 1-D matrices (coupled axis)
    def derivative_matrix(self):
        """(a,b) -> (a+1,b+1) coefficients; reference DifferentiateJacobi basis.py:701-718."""
        return (jacobi.differentiation_matrix(self.size, self.a, self.b) / self.COV.stretch).tocsr()

    def conversion_matrix(self, out):
        """(a,b) -> (out.a,out.b); reference ConvertJacobi basis.py:664-679."""
        return jacobi.conversion_matrix(self.size, self.a, self.b, out.a, out.b)

    def embed_constant_vector(self):
        """Constant 1 -> coefficients (reference ConvertConstantJacobi basis.py:682-698)."""
        v = np.zeros((self.size, 1)); v[0, 0] = 1 / self.constant_mode_value
        return sparse.csr_matrix(v)

    def integration_vector(self):
        return sparse.csr_matrix((jacobi.integration_vector(self.size, self.a, self.b) * self.COV.stretch)[None, :])

    def average_vector(self):
        return sparse.csr_matrix((jacobi.integration_vector(self.size, self.a, self.b) / 2)[None, :])

    def interpolation_vector(self, position):
        x = float(self.COV.native_coord(position))
        return sparse.csr_matrix(jacobi.interpolation_vector(self.size, self.a, self.b, x)[None, :])

    def valid_coeff_mask(self):
        return np.ones(self.size, dtype=bool)


def Legendre(*args, **kw):
    return Jacobi(*args, a=0, b=0, **kw)


def Ultraspherical(*args, alpha, alpha0=None, **kw):
    if alpha0 is None:
        alpha0 = alpha
    return Jacobi(*args, a=alpha - 0.5, b=alpha - 0.5, a0=alpha0 - 0.5, b0=alpha0 - 0.5, **kw)


def ChebyshevT(*args, **kw):
    return Ultraspherical(*args, alpha=0, **kw)


def ChebyshevU(*args, **kw):
    return Ultraspherical(*args, alpha=1, **kw)


def ChebyshevV(*args, **kw):
    return Ultraspherical(*args, alpha=2, **kw)


Chebyshev = ChebyshevT


def Fourier(*args, dtype=None, **kw):
    """Fourier basis factory (reference core/basis.py:1293-1300): real or complex by dtype."""
    if dtype is None:
        raise ValueError("dtype must be specified")
    if np.issubdtype(np.dtype(dtype), np.complexfloating):
        return ComplexFourier(*args, **kw)
    return RealFourier(*args, **kw)
