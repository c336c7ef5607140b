"""Field: data container with a device-resident buffer and an on-demand host mirror.

Mirrors the reference Field API used by the stock IVP scripts (core/field.py:345-1043): item access
`f['g']`, `f['c']`, `f[('g', scale)]`, `f['g'] = ...`, `.data`, `change_scales`, `change_layout`,
`fill_random` (bit-identical global stream, tools/random_arrays.py:7-55), `copy`.
SURVEY.md H7: user scripts mutate the *returned* numpy view (`b['g'] *= ...`, `u['g'][0][mask] = ...`), so
item access hands out the host mirror and marks it authoritative; the device copy is refreshed lazily
before the next device operation (solver.step / layout change).
"""
import numbers
import numpy as np
from .operators import Operand


class Field(Operand):
    def __init__(self, dist, bases=None, name=None, tensorsig=None, dtype=None):
        self.dist = dist
        self.name = name
        self.bases = dist.bases_by_axis(bases)
        self.tensorsig = tuple(tensorsig) if tensorsig else ()
        self.dtype = np.dtype(dtype if dtype is not None else dist.dtype).type
        self.args = []
        self.layout = 'c'
        self.scales = tuple(1 for _ in range(dist.dim))
        self._host = np.zeros(self.local_shape('c', self.scales), dtype=self.dtype)
        self._dev = None
        self._fresh = 'host'
        dist._fields.append(self)

    def __repr__(self):
        return f"<Field {self.name}>"

    def atoms(self):
        return [self]

    def unique_bases(self):
        out = []
        for b in self.bases:
            if b is not None and not any(b is o for o in out):
                out.append(b)
        return tuple(out)

    # ---- shapes --------------------------------------------------------------------------------------
    @property
    def dealias(self):
        return tuple(1 if b is None else b.axis_dealias(self.dist.basis_subaxis(b, ax)) for ax, b in enumerate(self.bases))

    @property
    def is_real(self):
        return np.issubdtype(self.dtype, np.floating)

    def global_shape(self, layout, scales):
        shp = []
        for ax, b in enumerate(self.bases):
            if b is None:
                shp.append(1)
            elif layout == 'c':
                shp.append(b.axis_size(self.dist.basis_subaxis(b, ax)))
            else:
                shp.append(b.axis_grid_size(scales[ax], self.dist.basis_subaxis(b, ax)))
        return tuple(shp)

    def local_slices(self, layout, scales):
        out = []
        for ax, b in enumerate(self.bases):
            if layout == 'c':
                out.append(self.dist.coeff_local_slice(ax, b))
            else:
                out.append(self.dist.grid_local_slice(ax, b, scales[ax]))
        return tuple(out)

    def local_shape(self, layout, scales):
        sl = self.local_slices(layout, scales)
        return self.tshape + tuple(s.stop - s.start for s in sl)

    # ---- host / device mirrors ----------------------------------------------------------------------
    @property
    def data(self):
        """Host view in the current layout (authoritative after access)."""
        if self._fresh == 'dev':
            self._host = self._dev.detach().cpu().numpy().reshape(self.local_shape(self.layout, self.scales)).copy()
        self._fresh = 'host'
        return self._host

    @data.setter
    def data(self, value):
        self.data[...] = value

    def device_data(self):
        """Device tensor in the current layout (uploads the host mirror if it is newer)."""
        import torch
        from .lib import compute_device
        dev = compute_device()
        if self._fresh == 'host' or self._dev is None:
            self._dev = torch.from_numpy(np.ascontiguousarray(self._host)).to(dev)
            self._fresh = 'both'
        return self._dev

    def set_device_data(self, tensor, layout, scales=None):
        """Adopt a device tensor as the authoritative data in the given layout."""
        self.layout = layout
        if scales is not None:
            self.scales = tuple(scales)
        self._dev = tensor
        self._fresh = 'dev'

    def preset_layout(self, layout):
        layout = self._layout_name(layout)
        if layout != self.layout or self._host.shape != self.local_shape(layout, self.scales):
            self.layout = layout
            self._host = np.zeros(self.local_shape(layout, self.scales), dtype=self.dtype)
            self._dev = None
            self._fresh = 'host'

    def preset_scales(self, scales):
        scales = self.dist.remedy_scales(scales)
        if scales != self.scales:
            self.scales = scales
            self._host = np.zeros(self.local_shape(self.layout, scales), dtype=self.dtype)
            self._dev = None
            self._fresh = 'host'

    @staticmethod
    def _layout_name(layout):
        if layout in ('c', 'coeff'):
            return 'c'
        if layout in ('g', 'grid'):
            return 'g'
        raise ValueError(f"Unknown layout: {layout}")

    def change_scales(self, scales):
        scales = self.dist.remedy_scales(scales)
        if scales == self.scales:
            return
        if self.layout == 'g':
            self.change_layout('c')
        # coefficient data is scale independent
        self.scales = scales

    def change_layout(self, layout):
        layout = self._layout_name(layout)
        if layout == self.layout:
            return
        from .transforms import transform_field
        transform_field(self, layout)

    require_layout = change_layout

    def require_coeff_space(self):
        self.change_layout('c')

    def require_grid_space(self, scales=None):
        if scales is not None:
            self.change_scales(scales)
        self.change_layout('g')

    def __getitem__(self, key):
        if isinstance(key, tuple):
            layout, scales = key
            self.change_scales(scales)
        else:
            layout = key
        self.change_layout(layout)
        return self.data

    def __setitem__(self, key, value):
        if isinstance(key, tuple):
            layout, scales = key
            self.preset_scales(scales)
        else:
            layout = key
        self.preset_layout(layout)
        np.copyto(self.data, value)

    def copy_device_to_grid(self, scales=None):
        """Grid values (dealias scales by default) of this field as a NEW device tensor; the field itself is left untouched."""
        tmp = Field(self.dist, bases=self.unique_bases(), tensorsig=self.tensorsig, dtype=self.dtype)
        self.dist._fields.pop()
        if self.layout != 'c':
            self.change_layout('c')
        tmp.scales = self.dealias if scales is None else self.dist.remedy_scales(scales)
        tmp.set_device_data(self.device_data().clone(), 'c')
        tmp.change_layout('g')
        return tmp.device_data()

    def copy(self):
        out = Field(self.dist, bases=self.unique_bases(), tensorsig=self.tensorsig, dtype=self.dtype)
        out.preset_scales(self.scales)
        out.preset_layout(self.layout)
        np.copyto(out.data, self.data)
        return out

    def fill_random(self, layout=None, scales=None, seed=None, chunk_size=2**20, distribution='standard_normal', **kw):
        """Reproduce the reference's mesh-independent global random stream (field.py:898-943)."""
        init_layout = self.layout
        if scales is not None:
            self.preset_scales(scales)
            if layout is None:
                self.preset_layout(init_layout)
        if layout is not None:
            self.preset_layout(layout)
        gshape = self.tshape + self.global_shape(self.layout, self.scales)
        if not self.is_real:
            gshape = gshape + (2,)
        total = int(np.prod(gshape))
        rng = np.random.default_rng(seed)
        draw = getattr(rng, distribution)
        csz = min(total, chunk_size) if total else chunk_size
        chunks = []
        done = 0
        while done < total:
            chunks.append(draw(size=csz, **kw))
            done += csz
        flat = np.concatenate(chunks)[:total] if chunks else np.zeros(0)
        gdata = flat.reshape(gshape)
        sl = tuple(slice(None) for _ in self.tensorsig) + self.local_slices(self.layout, self.scales)
        if self.is_real:
            self.data[...] = gdata[sl]
        else:
            loc = gdata[sl + (slice(None),)]
            self.data.real[...] = loc[..., 0]
            self.data.imag[...] = loc[..., 1]

    # ---- global data, norms and spectral filters (reference core/field.py:624-630, 746-876, 945-986) --------------------------
    def evaluate(self):
        return self

    def _full_slices(self):
        return tuple(slice(None) for _ in self.tensorsig) + self.local_slices(self.layout, self.scales)

    def set_global_data(self, global_data):
        np.copyto(self.data, np.asarray(global_data)[self._full_slices()])

    def load_from_global_grid_data(self, global_data, pre_slices=()):
        self.preset_layout('g')
        np.copyto(self.data, np.asarray(global_data)[tuple(pre_slices) + self._full_slices()])

    def load_from_global_coeff_data(self, global_data, pre_slices=()):
        self.preset_layout('c')
        np.copyto(self.data, np.asarray(global_data)[tuple(pre_slices) + self._full_slices()])

    def _host_allreduce(self, array, op):
        """All-reduce of a host array over the ranks of the mesh (through the device for NCCL)."""
        import torch, torch.distributed as td
        dev = 'cuda' if td.get_backend() == 'nccl' else 'cpu'
        t = torch.from_numpy(np.ascontiguousarray(array)).to(dev)
        if t.is_complex():
            t = torch.view_as_real(t).contiguous()
        td.all_reduce(t, op=getattr(td.ReduceOp, op))
        out = t.cpu().numpy()
        return out.view(array.dtype).reshape(array.shape) if np.iscomplexobj(array) else out

    def allgather_data(self, layout=None):
        """The global data on every rank (reference field.py:782-800)."""
        if layout is not None:
            self.change_layout(layout)
        if self.dist.size == 1:
            return self.data.copy()
        send = np.zeros(self.tshape + self.global_shape(self.layout, self.scales), dtype=self.dtype)
        send[self._full_slices()] = self.data
        return self._host_allreduce(send, 'SUM')

    def gather_data(self, root=0, layout=None):
        """The global data on rank `root`, None elsewhere (reference field.py:802-818)."""
        data = self.allgather_data(layout)
        return data if self.dist.rank == root else None

    def allreduce_data_norm(self, layout=None, order=2):
        if layout is not None:
            self.change_layout(layout)
        a = np.abs(self.data)
        if order == np.inf:
            norm = np.array([a.max() if a.size else 0.0])
            return float(self._host_allreduce(norm, 'MAX')[0]) if self.dist.size > 1 else float(norm[0])
        norm = np.array([np.sum(a ** order)])
        if self.dist.size > 1:
            norm = self._host_allreduce(norm, 'SUM')
        return float(norm[0] ** (1 / order))

    def allreduce_data_max(self, layout=None):
        return self.allreduce_data_norm(layout=layout, order=np.inf)

    def allreduce_L2_norm(self, normalize_volume=True):
        """sqrt of the volume average (or integral) of |f|^2 (reference field.py:844-862), real fields up to rank 1."""
        from . import operators as ops
        if not self.is_real or len(self.tensorsig) > 1:
            raise NotImplementedError("L2 norms of complex or rank-2 fields")
        inner = (self * self) if not self.tensorsig else ops.DotProduct(self, self)
        red = ops.Integrate(inner, average=bool(normalize_volume)).evaluate()
        return float(red.allreduce_data_max(layout='g')) ** 0.5

    def normalize(self, normalize_volume=True):
        norm = self.allreduce_L2_norm(normalize_volume=normalize_volume)
        self.data[...] = self.data / norm

    def low_pass_filter(self, shape=None, scales=None):
        """Zero the modes above the given relative scales: a round trip through a grid of that size (reference field.py:945-967)."""
        original = self.scales
        if shape is not None:
            if scales is not None:
                raise ValueError("Specify either shape or scales.")
            scales = tuple(np.array(shape) / np.array(self.global_shape('g', self.dist.remedy_scales(1))))
        self.change_scales(scales)
        self.change_layout('g')
        self.change_scales(original)

    def high_pass_filter(self, shape=None, scales=None):
        orig = np.array(self['c'])
        self.low_pass_filter(shape=shape, scales=scales)
        filt = np.array(self['c'])
        self['c'] = orig - filt
