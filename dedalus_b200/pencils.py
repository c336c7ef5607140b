"""Pencil systems: template assembly, splitting into independent components, static ordering and the
integer "programs" executed by the batched CUDA kernels (csrc/pencil.cu).

Reference behaviour being replaced (host side, setup time):
  * core/subsystems.py:34-53, 72-81   one Subproblem per (kx, ky) group, assembled in a Python loop
  * core/subsystems.py:497-602        expression matrices -> COO -> valid-mode filter -> permutations
  * libraries/matsolvers.py:126-183   SuperLU(COLAMD) factorisation of each pencil's  a0*M + b0*L
B200-first redesign (DESIGN.md section "pencil systems"):
  * pencil matrices are polynomials in the separable wavenumbers:  A(k) = sum_m k^m T_m ; only the
    templates T_m are assembled on the host (once per *class* of pencils: which wavenumbers are zero);
  * every class splits into independent components (parity blocks); each (class, component) is a *batch*
    of structurally identical systems, stored structure-of-arrays ([entry][system]) so that one GPU thread
    owns one system and every load/store is coalesced across systems;
  * one static row/column ordering per batch (max-product matching on a representative pencil, the
    MC64 + static-pivoting idea of distributed sparse solvers) lets all systems share one symbolic LU:
    the fill pattern, the elimination schedule and the triangular-solve schedule are integer programs
    built here, executed verbatim by every thread.
"""
import itertools
import numpy as np
from scipy import sparse
from scipy.sparse import csgraph
from scipy.optimize import linear_sum_assignment
from .operators import linear_map, Operand
from .basis import RealFourier, ComplexFourier, Jacobi


class Slot:
    """One line of coefficients along the coupled axis: (field/equation index, tensor comp, parity indices)."""
    __slots__ = ("owner", "comp", "par", "size", "has_last", "valid")

    def __init__(self, owner, comp, par, size, has_last, valid):
        self.owner, self.comp, self.par, self.size, self.has_last, self.valid = owner, comp, par, size, has_last, valid


def _enumerate_slots(items, sep_axes, zero_axes, last_axis):
    """items: list of (tensor shape, bases).  Returns list of Slot (natural reference ordering:
    owner-major, then comp, then separable parities, then last-axis index) and natural offsets."""
    slots, offsets = [], []
    off = 0
    for owner, (tshape, bases) in enumerate(items):
        ncomp = int(np.prod(tshape, dtype=int))
        gsz = []
        present = True
        for ax in sep_axes:
            b = bases[ax]
            if b is None:
                gsz.append(1)
                if ax not in zero_axes:
                    present = False
            else:
                gsz.append(b.group_size)
        if not present:
            continue
        last = bases[last_axis]
        nlast = 1 if last is None else last.size
        for comp in range(ncomp):
            for par in itertools.product(*[range(g) for g in gsz]):
                valid = True
                for ax, p in zip(sep_axes, par):
                    b = bases[ax]
                    if isinstance(b, RealFourier) and ax in zero_axes and p == 1:
                        valid = False     # -sin(0 x) slot (reference basis.py:1123-1134)
                slots.append(Slot(owner, comp, par, nlast, last is not None, valid))
                offsets.append(off)
                off += nlast
    return slots, np.array(offsets, dtype=np.int64), off


class PencilClass:
    """All pencils sharing the same set of vanishing separable wavenumbers."""

    def __init__(self, zero_axes, groups):
        self.zero_axes = zero_axes
        self.groups = groups          # (S, nsep) integer group indices of member pencils
        self.templates = {}           # name -> {monomial tuple: csr (rows x cols, natural order incl. invalid)}


class PencilSystemBuilder:
    """Assemble templates for M and L of an IVP and split them into batches."""

    def __init__(self, problem, entry_cutoff=1e-12):
        self.problem = problem
        dist = problem.dist
        self.dist = dist
        self.dim = dist.dim
        self.last_axis = dist.dim - 1
        self.sep_axes = tuple(range(dist.dim - 1))
        self.entry_cutoff = entry_cutoff
        self.variables = problem.variables
        self.equations = problem.equations
        # separable bases (one per separable axis), taken from the variables
        self.sep_bases = []
        for ax in self.sep_axes:
            bs = {v.bases[ax] for v in self.variables if v.bases[ax] is not None}
            bs |= {eq['bases'][ax] for eq in self.equations if eq['bases'][ax] is not None}
            if len(bs) != 1:
                raise NotImplementedError("Exactly one Fourier basis per separable axis is supported.")
            b = bs.pop()
            if not isinstance(b, (RealFourier, ComplexFourier)):
                raise NotImplementedError("Only the last axis may be non-Fourier (coupled); leading axes must be Fourier.")
            self.sep_bases.append(b)
        # linear maps of every equation's LHS
        self.eq_terms = []
        for eq in self.equations:
            lm = linear_map(eq['LHS'], self.variables, self.last_axis)
            self.eq_terms.append(lm)
        self.classes = self._build_classes()

    # ---------------------------------------------------------------------------------------------
    def local_groups(self, ax):
        """Group indices along separable axis `ax` owned by this rank (axis 0 is block-distributed)."""
        b = self.sep_bases[ax]
        if isinstance(b, RealFourier):
            ng = b.size // 2
        else:
            ng = b.size
        if ax == 0 and self.dist.size > 1:
            s, e = self.dist.block_range(ng, self.dist.size, self.dist.rank)
            return np.arange(s, e)
        return np.arange(ng)

    def group_wavenumber(self, ax, g):
        b = self.sep_bases[ax]
        if isinstance(b, RealFourier):
            return np.asarray(g) / b.COV.stretch
        return b.wavenumbers[np.asarray(g)]

    def group_valid(self, ax, g):
        b = self.sep_bases[ax]
        if isinstance(b, ComplexFourier):
            return b.valid_coeff_mask()[np.asarray(g)]
        return np.ones(np.shape(g), dtype=bool)

    def _build_classes(self):
        per_axis = [self.local_groups(ax) for ax in self.sep_axes]
        classes = {}
        if not self.sep_axes:
            classes[()] = PencilClass((), np.zeros((1, 0), dtype=np.int64))
        else:
            mesh = np.stack(np.meshgrid(*per_axis, indexing='ij'), axis=-1).reshape(-1, len(self.sep_axes))
            ok = np.ones(len(mesh), dtype=bool)
            for i, ax in enumerate(self.sep_axes):
                ok &= self.group_valid(ax, mesh[:, i])
            mesh = mesh[ok]
            kz = np.stack([self.group_wavenumber(ax, mesh[:, i]) == 0 for i, ax in enumerate(self.sep_axes)], axis=1)
            for pat in sorted({tuple(r) for r in kz.tolist()}):
                sel = np.all(kz == np.array(pat)[None, :], axis=1)
                zero_axes = tuple(ax for ax, z in zip(self.sep_axes, pat) if z)
                classes[zero_axes] = PencilClass(zero_axes, mesh[sel])
        for cls in classes.values():
            self._assemble_class(cls)
        return classes

    # ---------------------------------------------------------------------------------------------
    def _assemble_class(self, cls):
        var_items = [(v.tshape, v.bases) for v in self.variables]
        eq_items = [(tuple(cs.dim for cs in eq['tensorsig']), eq['bases']) for eq in self.equations]
        cls.col_slots, cls.col_off, ncols = _enumerate_slots(var_items, self.sep_axes, cls.zero_axes, self.last_axis)
        cls.row_slots, cls.row_off, nrows = _enumerate_slots(eq_items, self.sep_axes, cls.zero_axes, self.last_axis)
        cls.shape = (nrows, ncols)
        cls.valid_rows = np.concatenate([np.full(s.size, s.valid) for s in cls.row_slots]) if cls.row_slots else np.zeros(0, bool)
        cls.valid_cols = np.concatenate([np.full(s.size, s.valid) for s in cls.col_slots]) if cls.col_slots else np.zeros(0, bool)
        # natural block offsets per owner
        def owner_ranges(slots, offs, total):
            r = {}
            for s, o in zip(slots, offs):
                lo, hi = r.get(s.owner, (o, o))
                r[s.owner] = (min(lo, o), max(hi, o + s.size))
            return r
        col_rng = owner_ranges(cls.col_slots, cls.col_off, ncols)
        row_rng = owner_ranges(cls.row_slots, cls.row_off, nrows)
        templates = {'M': {}, 'L': {}}
        for ie, lm in enumerate(self.eq_terms):
            if ie not in row_rng:
                continue
            r0, r1 = row_rng[ie]
            for var, terms in lm.items():
                iv = self.variables.index(var)
                if iv not in col_rng:
                    continue
                c0, c1 = col_rng[iv]
                for t in terms:
                    if t.tder > 1:
                        raise NotImplementedError("Only first-order time derivatives are supported.")
                    name = 'M' if t.tder == 1 else 'L'
                    # expand separable symbols into monomials
                    sym_lists = []
                    skip = False
                    for ax in self.sep_axes:
                        sym = t.ops[ax]
                        if ax in cls.zero_axes:
                            sym = {m: A for m, A in sym.items() if m == 0}
                        if not sym:
                            skip = True
                            break
                        sym_lists.append(list(sym.items()))
                    if skip:
                        continue
                    Z = sparse.csr_matrix(t.ops[self.last_axis])
                    for combo in itertools.product(*sym_lists):
                        mono = tuple(m for m, _ in combo)
                        K = sparse.csr_matrix(t.comp)
                        for _, A in combo:
                            K = sparse.kron(K, sparse.csr_matrix(A), format='csr')
                        blk = (t.coef * sparse.kron(K, Z, format='csr')).tocoo()
                        if blk.shape != (r1 - r0, c1 - c0):
                            raise RuntimeError(f"Template block shape mismatch: {blk.shape} vs {(r1 - r0, c1 - c0)}")
                        full = sparse.coo_matrix((blk.data, (blk.row + r0, blk.col + c0)), shape=cls.shape).tocsr()
                        d = templates[name]
                        d[mono] = d[mono] + full if mono in d else full
        for name in templates:
            for mono in list(templates[name]):
                T = templates[name][mono]
                T.sum_duplicates()
                if np.iscomplexobj(T.data) and np.all(T.data.imag == 0) and self.problem.dtype in (np.float64,):
                    T = T.real
                T.eliminate_zeros()
                templates[name][mono] = T.tocsr()
                if T.nnz == 0:
                    del templates[name][mono]
        cls.templates = templates
        if cls.valid_rows.sum() != cls.valid_cols.sum():
            raise ValueError(f"Non-square pencil system for class zero_axes={cls.zero_axes}: "
                             f"{int(cls.valid_rows.sum())} equations vs {int(cls.valid_cols.sum())} unknowns.")

    # ---------------------------------------------------------------------------------------------
    def monomial_values(self, cls, mono, groups=None):
        groups = cls.groups if groups is None else groups
        val = np.ones(len(groups), dtype=self.problem.dtype if np.issubdtype(self.problem.dtype, np.complexfloating) else np.float64)
        for i, ax in enumerate(self.sep_axes):
            if mono[i]:
                val = val * self.group_wavenumber(ax, groups[:, i]) ** mono[i]
        return val

    def class_matrix(self, cls, name, group, restrict=True):
        """Numerical pencil matrix (natural ordering) for one member `group` of the class."""
        group = np.asarray(group, dtype=np.int64).reshape(1, -1)
        A = sparse.csr_matrix(cls.shape, dtype=np.complex128 if any(np.iscomplexobj(T.data) for T in cls.templates[name].values()) else np.float64)
        for mono, T in cls.templates[name].items():
            A = A + self.monomial_values(cls, mono, group)[0] * T
        A = A.tocsr()
        # entry cutoff as in the reference (subsystems.py:536)
        A.data[np.abs(A.data) < self.entry_cutoff] = 0
        A.eliminate_zeros()
        if restrict:
            A = A[cls.valid_rows][:, cls.valid_cols]
        return A

    def find_class(self, group):
        group = tuple(int(g) for g in group)
        for cls in self.classes.values():
            if any(tuple(r) == group for r in cls.groups.tolist()):
                return cls
        raise KeyError(group)


# ------------------------------------------------------------------------------------------------------
# Batches: independent components of a class, ordering, symbolic LU and programs
# ------------------------------------------------------------------------------------------------------
def _pattern(cls):
    P = sparse.csr_matrix(cls.shape, dtype=np.int8)
    for name in ('M', 'L'):
        for T in cls.templates[name].values():
            P = P + (abs(T) > 0).astype(np.int8)
    P = P[cls.valid_rows][:, cls.valid_cols].tocsr()
    P.data[:] = 1
    return P


def split_components(cls):
    """Connected components of the bipartite row/column graph of the class pattern (valid entries only)."""
    P = _pattern(cls)
    nr, nc = P.shape
    B = sparse.bmat([[None, P], [P.T, None]], format='csr')
    ncomp, lab = csgraph.connected_components(B, directed=False)
    vr = np.nonzero(cls.valid_rows)[0]
    vc = np.nonzero(cls.valid_cols)[0]
    comps = []
    for c in range(ncomp):
        rows = vr[lab[:nr] == c]
        cols = vc[lab[nr:] == c]
        if len(rows) != len(cols):
            raise ValueError("Structurally non-square independent component in pencil system.")
        if len(rows):
            comps.append((rows, cols))
    return comps


def _slot_of(index, offsets):
    return np.searchsorted(offsets, index, side='right') - 1


class Batch:
    """Structurally identical systems:  one class component x all member pencils."""

    def __init__(self, builder, cls, rows, cols):
        self.builder, self.cls = builder, cls
        self.n = len(rows)
        self.groups = cls.groups
        self.S = len(cls.groups)
        # mode-major preliminary ordering: interior lines sorted by last-axis index, then slot; border last
        rs = _slot_of(rows, cls.row_off); cs = _slot_of(cols, cls.col_off)
        rmode = rows - cls.row_off[rs]; cmode = cols - cls.col_off[cs]
        rint = np.array([cls.row_slots[s].has_last for s in rs]); cint = np.array([cls.col_slots[s].has_last for s in cs])
        rorder = np.lexsort((rs, np.where(rint, rmode, 0), ~rint))
        corder = np.lexsort((cs, np.where(cint, cmode, 0), ~cint))
        self.rows = rows[rorder]          # natural (class) row index of solver row i
        self.cols0 = cols[corder]         # natural column index in preliminary order
        self.cols = None                  # after matching
        self.n_border_rows = int((~rint).sum())

    # -- numeric matrix of one member in the current ordering
    def matrix(self, name_or_coefs, group, cols=None):
        cols = self.cols if cols is None else cols
        b = self.builder
        if isinstance(name_or_coefs, str):
            A = b.class_matrix(self.cls, name_or_coefs, group, restrict=False)
        else:
            a0, b0 = name_or_coefs
            A = a0 * b.class_matrix(self.cls, 'M', group, restrict=False) + b0 * b.class_matrix(self.cls, 'L', group, restrict=False)
        return A.tocsr()[self.rows][:, cols].tocsr()

    def representative_group(self):
        g = self.groups
        if len(g) == 0:
            return None
        # member with median total wavenumber
        order = np.argsort(g.sum(axis=1), kind='stable')
        return g[order[len(order) // 2]]

    def compute_ordering(self, a0, b0):
        """Static column order: max-product transversal on a representative LHS (after row/col scaling)."""
        A = self.matrix((a0, b0), self.representative_group(), cols=self.cols0).toarray()
        absA = np.abs(A)
        r = absA.max(axis=1); r[r == 0] = 1
        Sc = absA / r[:, None]
        c = Sc.max(axis=0); c[c == 0] = 1
        Sc = Sc / c[None, :]
        with np.errstate(divide='ignore'):
            cost = np.where(Sc > 0, -np.log(Sc), 1e8)
        ri, ci = linear_sum_assignment(cost)
        if np.any(cost[ri, ci] >= 1e8):
            raise ValueError("Pencil system is structurally singular (no perfect matching).")
        self.cols = self.cols0[ci]
        return self.cols

    # -- symbolic factorisation on the union pattern -------------------------------------------------
    def structural_pattern(self):
        cls = self.cls
        P = sparse.csr_matrix(cls.shape, dtype=np.int8)
        for name in ('M', 'L'):
            for T in cls.templates[name].values():
                P = P + (abs(T) > 0).astype(np.int8)
        P = P.tocsr()[self.rows][:, self.cols]
        return P.toarray() > 0

    def symbolic_lu(self):
        """Fill pattern of LU without pivoting in the current ordering (boolean n x n)."""
        F = self.structural_pattern()
        n = self.n
        if not np.all(np.diag(F)):
            raise ValueError("Zero structural diagonal after matching.")
        for k in range(n - 1):
            below = np.nonzero(F[k + 1:, k])[0]
            if below.size:
                F[k + 1 + below, k + 1:] |= F[k, k + 1:]
        return F
