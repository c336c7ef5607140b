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
