"""Distributor: coordinate bookkeeping, process mesh and block pencil decomposition.

Mirrors reference core/distributor.py:44-310 (Distributor) and 311-517 (Layout) for the subset used by the
hot path: D-dimensional Cartesian domains on a 1-D process mesh (P ranks = P GPUs of one box), coefficient
space distributed along axis 0 in blocks of whole groups, grid space distributed along axis 1 after the
single x<->y transpose hop (reference _build_layouts 131-175 with R=1).
Ranks come from torch.distributed (NCCL on GPUs, gloo in CPU tests) instead of mpi4py.
"""
import math
import numpy as np
from .coords import Coordinate, CartesianCoordinates


def _dist_info():
    try:
        import torch.distributed as td
        if td.is_available() and td.is_initialized():
            return td.get_rank(), td.get_world_size()
    except Exception:
        pass
    return 0, 1


class Distributor:
    def __init__(self, coordsystems, comm=None, mesh=None, dtype=None, device=None):
        if not isinstance(coordsystems, (tuple, list)):
            coordsystems = (coordsystems,)
        self.coordsystems = tuple(coordsystems)
        self.coordsys = self.coordsystems[0] if len(self.coordsystems) == 1 else None
        self.coords = tuple(c for cs in self.coordsystems for c in cs.coords)
        for c in self.coords:
            c.dist = self
        self.dim = len(self.coords)
        self.dtype = np.dtype(dtype if dtype is not None else np.float64).type
        rank, size = _dist_info()
        if mesh is None:
            mesh = (size,) if (size > 1 and self.dim > 1) else ()
        mesh = tuple(int(m) for m in mesh if int(m) > 1)
        if len(mesh) > 1:
            raise NotImplementedError("Only 1-D process meshes are supported (one NVSwitch box).")
        if mesh and mesh[0] != size:
            raise ValueError(f"Process mesh {mesh} does not match world size {size}.")
        if len(mesh) >= self.dim:
            raise ValueError("Mesh must have lower dimension than distributor.")
        self.mesh = mesh
        self.rank, self.size = rank, size if mesh else 1
        self.device = device
        self._fields = []
        # the reference's scripts ask `dist.comm.rank == 0` / `dist.comm.size` (mpi4py communicator): the same two numbers here
        from types import SimpleNamespace
        self.comm = self.comm_cart = SimpleNamespace(rank=rank, size=size, Get_rank=lambda: rank, Get_size=lambda: size)

    # ---- coordinate bookkeeping --------------------------------------------------------------------
    def get_coord(self, name):
        for c in self.coords:
            if c.name == name:
                return c
        raise ValueError(f"Unknown coordinate name: {name}")

    def get_axis(self, coord):
        if isinstance(coord, str):
            coord = self.get_coord(coord)
        return self.coords.index(coord)

    def get_basis_axis(self, basis):
        return self.get_axis(basis.coord)

    def bases_by_axis(self, bases):
        if bases is None:
            bases = ()
        if not isinstance(bases, (tuple, list)):
            bases = (bases,)
        out = [None] * self.dim
        for b in bases:
            if b is None:
                continue
            ax = self.get_basis_axis(b)
            for sub in range(b.dim):            # a multi-dimensional basis (sphere) occupies dim consecutive axes
                if out[ax + sub] is not None:
                    raise ValueError("Overlapping bases specified.")
                out[ax + sub] = b
        return tuple(out)

    def basis_subaxis(self, basis, axis):
        """Index of `axis` inside `basis` (0 for one-dimensional bases)."""
        return axis - self.get_basis_axis(basis)

    # ---- field factories (reference distributor.py:213-235) ----------------------------------------
    def Field(self, *args, **kw):
        from .field import Field
        return Field(self, *args, **kw)

    ScalarField = Field

    def VectorField(self, coordsys, *args, **kw):
        from .field import Field
        return Field(self, *args, tensorsig=(coordsys,), **kw)

    def TensorField(self, tensorsig, *args, **kw):
        from .field import Field
        if not isinstance(tensorsig, (tuple, list)):
            tensorsig = (tensorsig,)
        return Field(self, *args, tensorsig=tuple(tensorsig), **kw)

    # ---- block decomposition (reference Layout.local_chunks distributor.py:357-385) ---------------
    @staticmethod
    def block_range(nchunks, nranks, rank):
        block = -(-nchunks // nranks)
        start = min(block * rank, nchunks)
        return start, min(start + block, nchunks)

    def coeff_local_slice(self, axis, basis):
        """Local slice along `axis` in coefficient layout (axis 0 distributed in whole groups)."""
        size = 1 if basis is None else basis.axis_size(self.basis_subaxis(basis, axis))
        if axis != 0 or self.size == 1 or basis is None:
            return slice(0, size)
        g = basis.axis_group_size(0) if basis.dim > 1 else basis.group_size
        if basis.dim > 1 and (size // g) % self.size:
            raise ValueError(f"curvilinear bases need the {size // g} azimuthal pairs of the coefficient packing to divide evenly over {self.size} ranks")
        s, e = self.block_range(size // g, self.size, self.rank)
        return slice(s * g, e * g)

    def grid_local_slice(self, axis, basis, scale):
        """Local slice along `axis` in full grid layout (axis 1 distributed after the transpose hop)."""
        size = 1 if basis is None else basis.axis_grid_size(scale, self.basis_subaxis(basis, axis))
        if axis != 1 or self.size == 1 or basis is None:
            return slice(0, size)
        if basis.dim > 1 and size % self.size:
            raise ValueError(f"curvilinear bases need the {size} colatitude grid points to divide evenly over {self.size} ranks")
        s, e = self.block_range(size, self.size, self.rank)
        return slice(s, e)

    # ---- grids ---------------------------------------------------------------------------------------
    def local_grid(self, basis, scale=None):
        scale = 1 if scale is None else scale
        axis = self.get_basis_axis(basis)
        g = basis.global_grid(scale)
        g = g[self.grid_local_slice(axis, basis, scale)]
        shape = [1] * self.dim
        shape[axis] = g.size
        return g.reshape(shape)

    def local_grids(self, *bases, scales=None):
        scales = self.remedy_scales(scales)
        out = []
        for b in bases:
            ax = self.get_basis_axis(b)
            if b.dim > 1:       # one grid per axis of the basis (reference distributor.py:291-301)
                out.extend(b.local_grids(self, scales[ax:ax + b.dim]))
            else:
                out.append(self.local_grid(b, scales[ax]))
        return tuple(out)

    def remedy_scales(self, scales):
        if scales is None:
            scales = 1
        if np.isscalar(scales):
            scales = [scales] * self.dim
        if 0 in scales:
            raise ValueError("Scales must be nonzero.")
        return tuple(scales)
