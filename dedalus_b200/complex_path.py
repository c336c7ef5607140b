"""Right-hand sides of complex-dtype Cartesian IVPs (T3: ComplexFourier^n x Jacobi problems in complex128).

Reference: the same Evaluator walk as for real problems (core/evaluator.py:95-146) on complex128 fields, with
FFTWComplexFFT along the Fourier axes (core/transforms.py:302-330) and the Chebyshev transform applied to the interleaved real
and imaginary lines (libraries/fftw/fftw_wrappers.pyx:244-246).

Each right-hand side is lowered once (evaluator.lower) to  F = sum coef * prod d^alpha(field component)  and evaluated as
  1. backward transforms of the unique factors (complex tensors; derivatives ride in the transforms' load stages),
  2. ONE pointwise launch on real / imaginary PLANES: a complex product of k factors is 2^k real monomials per part,
  3. forward transforms of the outputs, written as planes into the equation arena the pencil gather reads
     (pencils.PencilSystemBuilder: complex unknowns are carried as two real ones).
Straightforward rather than fused: complex problems are not on the benchmark path.
"""
import itertools
import numbers
import numpy as np
from .evaluator import lower


class ComplexRHSPlan:
    def __init__(self, solver):
        import torch
        self.solver = solver
        problem = solver.problem
        dist = self.dist = problem.dist
        dev = self.device = solver.device
        arena = solver.eq_arena
        self.static_entries = []
        inputs, self.input_keys, outputs = {}, [], []
        for ie, eq in enumerate(problem.equations):
            rhs = eq['RHS']
            if isinstance(rhs, numbers.Number):
                if rhs != 0:
                    self._add_constant(ie, eq, complex(rhs))
                continue
            for comp, terms in lower(rhs).items():
                const = sum(c for c, f in terms if len(f) == 0)
                terms = [(c, f) for c, f in terms if len(f) > 0]
                if const != 0:
                    self._add_constant(ie, eq, complex(const), comp)
                if not terms:
                    continue
                if any(b is None for b in eq['bases']):
                    raise NotImplementedError("Field-dependent RHS of a lower-dimensional equation is not supported yet.")
                tl = []
                for coef, facs in terms:
                    idxs = []
                    for (f, c, dv) in facs:
                        if any(b is None for b in f.bases):
                            raise NotImplementedError("RHS factors must have bases along every axis.")
                        key = (id(f), c, dv)
                        if key not in inputs:
                            inputs[key] = len(self.input_keys)
                            self.input_keys.append((f, c, dv))
                        idxs.append(inputs[key])
                    tl.append((complex(coef), idxs))
                outputs.append((ie, comp, tl))
        self.outputs = outputs
        self.n_in, self.n_out = len(self.input_keys), len(outputs)
        if self.n_out == 0:
            return
        f0 = self.input_keys[0][0]
        self.bases = f0.bases
        self.dealias = tuple(b.dealias[0] for b in self.bases)
        self.gshape = tuple(b.grid_size(s) for b, s in zip(self.bases, self.dealias))
        self.npoints = int(np.prod(self.gshape))
        # ---- real program on planes: input plane 2 i + part, output plane 2 o + part
        term_ptr, coef, fac_ptr, fac = [0], [], [0], []
        for ie, comp, tl in outputs:
            for part in (0, 1):
                for c, idxs in tl:
                    # c * prod_j (x_j + i y_j): choose the imaginary part of the factors in `sel`; i^|sel| = (1, i, -1, -i)
                    for sel in itertools.product((0, 1), repeat=len(idxs)):
                        w = c * (1j ** sum(sel))
                        v = w.real if part == 0 else w.imag
                        if v == 0:
                            continue
                        coef.append(float(v))
                        fac.extend(2 * i + s for i, s in zip(idxs, sel))
                        fac_ptr.append(len(fac))
                term_ptr.append(len(coef))
        i32 = lambda a: torch.tensor(a, dtype=torch.int32, device=dev)
        self.term_ptr, self.fac_ptr, self.fac = i32(term_ptr), i32(fac_ptr), i32(fac if fac else [0])
        self.coef = torch.tensor(coef if coef else [0.0], dtype=torch.float64, device=dev)
        self.nfac = len(fac)
        self.planes_in = torch.empty((2 * self.n_in,) + self.gshape, dtype=torch.float64, device=dev)
        self.planes_out = torch.empty((2 * self.n_out,) + self.gshape, dtype=torch.float64, device=dev)
        self.out_views = []
        for ie, comp, tl in outputs:
            tsh, shp = arena.shapes[ie]
            n = int(np.prod(shp))
            ncomp = int(np.prod(tsh[1:], dtype=int))
            base = arena.offsets[ie]
            self.out_views.append((base + comp * n, base + (ncomp + comp) * n, shp, problem.equations[ie]['bases']))

    def _add_constant(self, ie, eq, value, comp=0):
        """Constant RHS = coefficient of the constant mode of the equation's domain, real and imaginary plane."""
        from .basis import Jacobi
        arena = self.solver.eq_arena
        tsh, shp = arena.shapes[ie]
        factor = 1.0
        for ax, b in enumerate(eq['bases']):
            if isinstance(b, Jacobi):
                factor /= b.constant_mode_value
        n = int(np.prod(shp))
        ncomp = int(np.prod(tsh[1:], dtype=int))
        self.static_entries.append((arena.offsets[ie] + comp * n, value.real * factor))
        self.static_entries.append((arena.offsets[ie] + (ncomp + comp) * n, value.imag * factor))

    def set_static(self, arena_tensor):
        for off, val in self.static_entries:
            arena_tensor[off] = val

    @staticmethod
    def _comp_index(f, comp):
        if not f.tensorsig:
            return ()
        return tuple(int(i) for i in np.unravel_index(comp, f.tshape))

    def evaluate(self, eq_t):
        import torch
        from .transforms import cached_plan
        from .lib import get_lib, current_stream
        from .solvers import Timed
        if self.n_out == 0:
            return
        dim = self.dist.dim
        prof = self.solver.prof
        for f in {id(k[0]): k[0] for k in self.input_keys}.values():
            if f.layout != 'c':
                f.change_layout('c')
        # ---- 1. backward transforms, last axis first
        for i, (f, c, dv) in enumerate(self.input_keys):
            cur = f.device_data()[self._comp_index(f, c)].contiguous()
            for ax in range(dim - 1, -1, -1):
                plan = cached_plan(f.bases[ax], self.dealias[ax])
                shp = list(cur.shape); shp[ax] = plan.N
                out = torch.empty(shp, dtype=cur.dtype, device=cur.device)
                with Timed(prof, f"transform_bwd_axis{ax}", 16 * (cur.numel() + out.numel())):
                    plan.backward(cur, out, ax, deriv=dv[ax]) if dv[ax] else plan.backward(cur, out, ax)
                cur = out
            self.planes_in[2 * i].copy_(cur.real); self.planes_in[2 * i + 1].copy_(cur.imag)
        # ---- 2. products
        with Timed(prof, "pointwise", 8 * self.npoints * 2 * (self.n_in + self.n_out)):
            get_lib().call("db_pointwise", self.planes_in.data_ptr(), self.planes_out.data_ptr(), self.npoints, 2 * self.n_in, 2 * self.n_out,
                           self.term_ptr.data_ptr(), self.coef.data_ptr(), self.fac_ptr.data_ptr(), self.fac.data_ptr(), self.nfac, current_stream())
        # ---- 3. forward transforms into the equation arena (planes)
        for o, (off_re, off_im, shp, bases) in enumerate(self.out_views):
            cur = torch.complex(self.planes_out[2 * o], self.planes_out[2 * o + 1])
            for ax in range(dim):
                plan = cached_plan(bases[ax], self.dealias[ax])
                s = list(cur.shape); s[ax] = plan.M
                out = torch.empty(s, dtype=cur.dtype, device=cur.device)
                with Timed(prof, f"transform_fwd_axis{ax}", 16 * (cur.numel() + out.numel())):
                    plan.forward(cur, out, ax)
                cur = out
            n = cur.numel()
            eq_t[off_re:off_re + n].copy_(cur.real.reshape(-1)); eq_t[off_im:off_im + n].copy_(cur.imag.reshape(-1))
