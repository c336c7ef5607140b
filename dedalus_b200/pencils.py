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
        # complex-dtype problems (T3): every complex unknown is carried as two real ones.  The real / imaginary PLANES of a
        # field are the leading tensor index of its arena entry, so all kernels see ordinary real coefficient arrays, and a
        # complex operator block B becomes [[Re B, -Im B], [Im B, Re B]] (the reference solves complex pencils in complex
        # arithmetic, libraries/matsolvers.py:126-183 on complex128 matrices).
        self.complex = bool(np.issubdtype(problem.dtype, np.complexfloating))
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
        var_items, eq_items = self.arena_items()
        cls.col_slots, cls.col_off, ncols = _enumerate_slots(var_items, self.sep_axes, cls.zero_axes, self.last_axis)
        cls.row_slots, cls.row_off, nrows = _enumerate_slots(eq_items, self.sep_axes, cls.zero_axes, self.last_axis)
        cls.shape = (nrows, ncols)
        cls.valid_rows = np.concatenate([np.full(s.size, s.valid) for s in cls.row_slots]) if cls.row_slots else np.zeros(0, bool)
        cls.valid_cols = np.concatenate([np.full(s.size, s.valid) for s in cls.col_slots]) if cls.col_slots else np.zeros(0, bool)
        # invalid modes along a coupled Fourier axis (-sin(0x) slot, Nyquist): reference basis.py:1123-1134
        for slots, offs, items, valid in ((cls.row_slots, cls.row_off, eq_items, cls.valid_rows),
                                          (cls.col_slots, cls.col_off, var_items, cls.valid_cols)):
            for s, o in zip(slots, offs):
                lb = items[s.owner][1][self.last_axis]
                if s.has_last and isinstance(lb, (RealFourier, ComplexFourier)):
                    valid[o:o + s.size] &= lb.valid_coeff_mask()
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
                        blk = (t.coef * sparse.kron(K, Z, format='csr'))
                        if self.complex:
                            blk = sparse.bmat([[blk.real, -blk.imag], [blk.imag, blk.real]], format='csr')
                        blk = blk.tocoo()
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
                # entry cutoff as in the reference (subsystems.py:536: |a| < entry_cutoff dropped after assembly), applied
                # to the template with the largest monomial value of the class: rounding-noise entries of sparse
                # products (1e-16) would otherwise bloat the structural pattern and the LU fill
                scale = float(np.max(np.abs(self.monomial_values(cls, mono)))) if len(cls.groups) else 1.0
                T.data[np.abs(T.data) * max(scale, 1e-300) < self.entry_cutoff] = 0
                T.eliminate_zeros()
                templates[name][mono] = T.tocsr()
                if T.nnz == 0:
                    del templates[name][mono]
        cls.templates = templates
        if cls.valid_rows.sum() != cls.valid_cols.sum():
            raise ValueError(f"Non-square pencil system for class zero_axes={cls.zero_axes}: "
                             f"{int(cls.valid_rows.sum())} equations vs {int(cls.valid_cols.sum())} unknowns.")

    def arena_items(self):
        """(tensor shape, bases) of every variable / equation as stored in the solver's arenas (complex: leading re / im index)."""
        lead = (2,) if self.complex else ()
        var_items = [(lead + tuple(v.tshape), v.bases) for v in self.variables]
        eq_items = [(lead + tuple(cs.dim for cs in eq['tensorsig']), eq['bases']) for eq in self.equations]
        return var_items, eq_items

    # ---------------------------------------------------------------------------------------------
    def monomial_values(self, cls, mono, groups=None):
        groups = cls.groups if groups is None else groups
        val = np.ones(len(groups), dtype=np.float64)
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
        self.seq = None                   # pivot sequence: cols = cols0[seq]
        self.n_border_rows = int((~rint).sum())
        # sign-equivalent components merged into this batch (merge_sign_equivalent): member c has the matrices
        #   A_c = D1_c A_0 D2_c   (D = diag(+-1)),  A_0 = this batch's own (canonical) component,
        # so with the vectors of member c stored as  D2_c x_c  (column space) and  D1_c b_c  (row space) every member is
        # solved / multiplied with the SAME matrix: one LU factorisation per pencil serves all members.
        self.members = [dict(rows=self.rows, cols0=self.cols0, d1=np.ones(self.n), d2=np.ones(self.n))]

    @property
    def R(self):
        return len(self.members)

    def describe(self, side, member=0):
        """(owner, tensor component, mode index, has_last) of every row / preliminary column position of a member."""
        cls = self.cls
        m = self.members[member]
        idx, offs, slots = (m['rows'], cls.row_off, cls.row_slots) if side == 'rows' else (m['cols0'], cls.col_off, cls.col_slots)
        sl = _slot_of(idx, offs)
        return np.array([(slots[q].owner, slots[q].comp, int(i - offs[q]), int(slots[q].has_last)) for q, i in zip(sl, idx)], dtype=np.int64)

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

    def representative_groups(self, nrep=6):
        """A few member pencils spanning the wavenumber range (corners, median, skewed)."""
        g = self.groups
        if len(g) <= nrep:
            return [r for r in g]
        picks = []
        tot = g.sum(axis=1)
        order = np.argsort(tot, kind='stable')
        picks += [order[0], order[-1], order[len(order) // 2]]
        for i in range(g.shape[1]):
            picks.append(np.lexsort((tot, -g[:, i]))[0])       # largest along axis i, smallest elsewhere
            picks.append(np.lexsort((-tot, g[:, i]))[0])       # smallest along axis i, largest elsewhere
        seen, out = set(), []
        for p in picks:
            if int(p) not in seen:
                seen.add(int(p)); out.append(g[int(p)])
        return out

    # Default pivot threshold.  The Chebyshev tau Helmholtz rows  (1 + eps k^2) S - eps D2  have the entries [a, b, a] with
    # |a| / |b| = 1 / (2 + 16 j^2 eps) < 1/2 (S: conversion T -> C^(2), D2: second derivative, j the mode): eliminating on
    # `a` runs the three-term recurrence in its GROWING direction (roots t, 1/t of t^2 - (2 + 16 j^2 eps) t + 1 = 0; growth
    # prod_j t_j ~ 1e22 at Nz = 256, dt = 2.5e-3 -- round 1's threshold 0.1 + Markowitz tie-break picked it), on `b` in the
    # decaying one.  Any threshold > 1/2 therefore takes `b`; 0.6 keeps the Markowitz freedom elsewhere (LU fill 18.5 / row
    # at Nz = 256 against 17.1 at 0.1 and 19.7 at 1.0).  The order is VERIFIED on the device for every system after each
    # factorisation (solvers.BatchSet.verify) and recomputed with threshold 1.0 + the offending members if it fails.
    DEFAULT_THRESHOLD = 0.6

    def compute_ordering(self, a0, b0, threshold=None, extra_groups=(), extra_lhs=()):
        """Static column (pivot) order shared by every system of the batch.

        Rows are processed in mode-major order (dense boundary rows last).  The pivot column of each row is
        chosen by *joint* threshold pivoting over several representative member pencils: a candidate's score
        is the minimum over representatives of |a_ij| / max_j |a_ij| (after the eliminations so far), so exact
        cancellations (e.g. the proportional pressure-gradient rows of RB) and wavenumber-regime changes are
        seen before the order is frozen.  Among candidates within `threshold` of the best score the one
        with the fewest remaining column entries (Markowitz) and lowest index is taken, which keeps the fill
        local.  This is what partial pivoting of A^T (reference matsolvers.py:179-183, SuperLU on A^T) does
        per pencil; here it is done once per batch and the GPU factorisation needs no pivot search.
        `extra_groups`: further member pencils to include as representatives (systems that failed verification);
        `extra_lhs`: further (a0, b0) pairs every representative is also taken at (timestep changes).
        """
        if threshold is None:
            # DB_PIVOT_THRESHOLD: diagnostic override (tests use 0.1, round 1's unstable choice, to exercise the
            # verify -> re-order path)
            import os
            threshold = float(os.environ.get("DB_PIVOT_THRESHOLD", self.DEFAULT_THRESHOLD))
        reps = list(self.representative_groups()) + [np.asarray(g) for g in extra_groups]
        lhs = [(a0, b0)] + [tuple(x) for x in extra_lhs]
        mats = np.stack([self.matrix(ab, g, cols=self.cols0).toarray() for ab in lhs for g in reps], axis=0)
        R, n, _ = mats.shape
        used = np.zeros(n, dtype=bool)
        seq = np.zeros(n, dtype=np.int64)
        tiny = 1e-13
        for i in range(n):
            rowabs = np.abs(mats[:, i, :])
            rowabs[:, used] = 0
            rmax = rowabs.max(axis=1)
            if np.any(rmax == 0):
                raise ValueError("Pencil system is singular for a representative wavenumber (zero row during ordering).")
            rel = rowabs / rmax[:, None]
            score = rel.min(axis=0)
            best = score.max()
            if best < 1e-8:
                raise ValueError("No jointly acceptable pivot found while ordering the pencil system.")
            cand = np.nonzero(score >= threshold * best)[0]
            if len(cand) > 1:
                below = (np.abs(mats[0, i + 1:, :][:, cand]) > tiny * rmax[0]).sum(axis=0)
                cand = cand[np.lexsort((cand, -score[cand], below))]
            j = int(cand[0])
            seq[i] = j
            used[j] = True
            # eliminate column j from the remaining rows (all representatives)
            piv = mats[:, i, j]
            col = mats[:, i + 1:, j]
            rows = np.nonzero(np.any(col != 0, axis=0))[0]
            if rows.size:
                f = col[:, rows] / piv[:, None]
                mats[:, i + 1 + rows, :] -= f[:, :, None] * mats[:, i, :][:, None, :]
                mats[:, i + 1 + rows, j] = 0
        self.cols = self.cols0[seq]
        self.seq = seq
        self.order_threshold = threshold
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
        for k in range(n - 1):
            if not F[k, k]:
                raise ValueError("Zero structural pivot in static ordering.")
            below = np.nonzero(F[k + 1:, k])[0]
            if below.size:
                F[k + 1 + below, k + 1:] |= F[k, k + 1:]
        return F


# ------------------------------------------------------------------------------------------------------
# Program compilation
# ------------------------------------------------------------------------------------------------------
class BatchProgram:
    """Integer programs + template values for one batch (all arrays are host numpy, uploaded by the solver)."""
    pass


def _permuted_templates(batch, name):
    """List of (mono, coo in solver ordering) for M or L."""
    out = []
    for mono, T in batch.cls.templates[name].items():
        P = T.tocsr()[batch.rows][:, batch.cols].tocoo()
        out.append((mono, P))
    return out


SOLVE_CTRL_WORDS = 44
SOLVE_DEEP_D = 3          # k_batches_solve_deep: gathers are issued this many chunks ahead (csrc/pencil.cu SOLVE_DEEP_D)
SOLVE_DEEP_RING = 16      # ... and the last 16 finished rows are kept in a shared-memory ring (SOLVE_DEEP_RRN)
SOLVE_PF_AHEAD = 12       # chunks between the prefetch of a row's start value and the chunk that enters the row


def solve_control_blocks(code, n_fwd, tile=64, CH=16):
    """Per-chunk control blocks of the branch-free solve kernels (csrc/pencil.cu), derived from the flat instruction stream:
    for each chunk of 16 entries
        goff[16]  element offset gathered before the chunk is consumed: the column of a multiply-accumulate entry, the
                  row ENTERED (its start value) at a row-boundary entry, 0 for padding (whose factor value is 0)
        foff[16]  element offset of the row LEFT at a row-boundary entry
        maskE     entries that leave a row (store), maskB entries that enter a row (take the gathered start value),
        maskF     entries whose gathered value was stored inside this very chunk (re-read right before use)
        maskF2    the same for a gather issued ONE CHUNK EARLIER (software-pipelined kernel variant): stored inside this
                  or the previous chunk
      deep-prefetch kernel (gathers issued SOLVE_DEEP_D chunks ahead, landing in shared memory):
        maskR     entries whose source row was finished after their gather was issued AND is still among the last
                  SOLVE_DEEP_RING finished rows: taken from the shared-memory ring of recent rows, slot rslot[j]
                  (= index of that row's store modulo the ring size; the kernel counts row ends the same way)
        maskG     such entries whose source is older than the ring: re-read from global memory right before use
        rslot     16 x 5 bits packed 6 per word
      prefetch (all kernels): pf[2] = element offsets of the first two rows ENTERED FOR THE FIRST TIME in chunk
        q + SOLVE_PF_AHEAD (-1: none).  A row's start value (the right-hand side in the forward sweep, the forward result
        in the backward sweep) was written long before it is needed and has left the L2 by then: without the prefetch every
        row costs one DRAM round trip on the critical path of its thread (ncu, round 2: that is what the time of the sweep
        was made of when there are few tiles per SM).
    Layout: int32 [nchunks][44] = goff, foff, maskE, maskB, maskF, maskF2, maskR, maskG, rslot[3], pf[2], spare."""
    SKIP = -2**31
    D, RRN = SOLVE_DEEP_D, SOLVE_DEEP_RING
    code = np.asarray(code, dtype=np.int64)
    nE = len(code)
    assert nE % CH == 0 and n_fwd % CH == 0
    out = np.zeros((nE // CH, SOLVE_CTRL_WORDS), dtype=np.int64)
    cur = -1
    stored_at = {}                         # element offset -> position of its most recent store
    store_idx = {}                         # element offset -> running index (count of row ends before it) of that store
    nstores = 0
    for e in range(nE):
        if e == n_fwd:
            cur = -1
        q, j = divmod(e, CH)
        c = int(code[e])
        if c == SKIP:
            continue

        def mark(src):
            pos = stored_at.get(src, -1)
            if pos >= q * CH:
                out[q, 34] |= 1 << j
            if pos >= (q - 1) * CH:
                out[q, 35] |= 1 << j
            if pos >= 0 and pos >= (q - D) * CH:          # stored after the deep kernel issued this entry's gather
                k = store_idx[src]
                if nstores - k <= RRN:                    # still in the ring (nstores = row ends before this entry)
                    out[q, 36] |= 1 << j
                    out[q, 38 + j // 6] |= (k % RRN) << (5 * (j % 6))
                else:
                    out[q, 37] |= 1 << j
        if c < 0:
            nxt = -1 - c
            out[q, j] = nxt
            out[q, 33] |= 1 << j
            mark(nxt)
            if cur >= 0:
                out[q, 16 + j] = cur
                out[q, 32] |= 1 << j
                stored_at[cur] = e
                store_idx[cur] = nstores
                nstores += 1
            cur = nxt
        else:
            out[q, j] = c
            mark(c)
    # prefetch words: first entries into rows, per chunk
    out[:, 41] = -1; out[:, 42] = -1
    first_entries = [[] for _ in range(nE // CH)]
    seen = set()
    for e in range(nE):
        if e == n_fwd:
            seen = set()
        c = int(code[e])
        if c != SKIP and c < 0:
            row = -1 - c
            if row not in seen:
                seen.add(row)
                first_entries[e // CH].append(row)
    for q in range(nE // CH):
        tgt = q + SOLVE_PF_AHEAD
        if tgt < nE // CH:
            for k, row in enumerate(first_entries[tgt][:2]):
                out[q, 41 + k] = row
    # (rows first entered in chunks 0 .. SOLVE_PF_AHEAD-1 are not prefetched; the kernel's prologue has just written them)
    return out.astype(np.int32)


def compile_batch(batch, a0, b0, dense=None):
    """Build ordering (if needed), symbolic LU and all programs for the LHS  a0*M + b0*L."""
    if batch.cols is None:
        batch.compute_ordering(a0, b0)
    b = batch.builder
    n = batch.n
    F = batch.symbolic_lu()
    prog = BatchProgram()
    prog.n, prog.S = n, batch.S
    # ---- solve stream (include/dedalus_b200.h).  Rows of each triangular solve are processed in LEVEL order of its
    #      dependency DAG (rows of one level are mutually independent), which keeps a row's inputs several rows behind it
    #      in the stream, so the 16 values a chunk needs can be gathered in one burst before it is consumed.  Codes:
    #        c >= 0          : acc -= LU[e] * x[c] ; c = column * TILE
    #        c <  0, != SKIP : leave the current row (store; backward: multiply by LU[e] = reciprocal pivot) and enter
    #                          row (-1 - c) / TILE with acc = x[row] ; the first entry of a section only enters a row
    #        DB_I_SKIP       : padding to a multiple of the chunk size
    SKIP, CH = -2**31, 16           # sections padded to the kernel's 16-entry chunks
    TILE = 64                        # DB_TILE: vectors / factors are stored tile-major, 64 systems per slab
    ld = ((batch.S + TILE - 1) // TILE) * TILE
    prog.ld, prog.tile = ld, TILE
    if (n + 1) * TILE >= 2**30:
        raise NotImplementedError("system too large for 30-bit vector offsets")
    lev_f = np.zeros(n, dtype=np.int64)
    for i in range(n):
        js = np.nonzero(F[i, :i])[0]
        lev_f[i] = 1 + (lev_f[js].max() if js.size else 0)
    lev_b = np.zeros(n, dtype=np.int64)
    for i in range(n - 1, -1, -1):
        js = i + 1 + np.nonzero(F[i, i + 1:])[0]
        lev_b[i] = 1 + (lev_b[js].max() if js.size else 0)
    import os
    if os.environ.get("DB_SOLVE_ORDER", "level") == "natural":
        order_f = np.arange(n); order_b = np.arange(n - 1, -1, -1)
    else:
        order_f = np.lexsort((np.arange(n), lev_f))
        order_b = np.lexsort((-np.arange(n), lev_b))
    prog.levels = (int(lev_f.max()), int(lev_b.max()))
    eid = -np.ones((n, n), dtype=np.int64)
    diag_eid = np.zeros(n, dtype=np.int32)
    code = []
    e = 0

    import os
    DENSE = int(dense if dense is not None else os.environ.get("DB_SOLVE_DENSE", 64))
    SEG = 15                          # forward rows with >= DENSE entries are visited in segments of SEG entries

    def visits_of(order, forward):
        """Sequence of (row, columns) visits.  Normally one visit per row.  A run of consecutive DENSE forward rows (the
        boundary rows, eliminated last: each sweeps the whole x vector) is interleaved segment by segment: every row
        takes the next SEG of the columns finished before the run, so the 15 x rows of a segment are re-used by all rows
        of the run while they are still cache-resident (one DRAM sweep of x for the run instead of one per row).  A row
        that is left and re-entered keeps its partial sum in x[row]: leaving a row always stores the accumulator and
        entering one loads x[row] (the forward sweep has no pivot scaling, so partial sums are exact)."""
        cols_of = (lambda i: np.nonzero(F[i, :i])[0]) if forward else (lambda i: i + 1 + np.nonzero(F[i, i + 1:])[0])
        out, idx = [], 0
        while idx < len(order):
            i = int(order[idx])
            js = cols_of(i)
            if not forward or js.size < DENSE:
                out.append((i, js)); idx += 1
                continue
            run = [i]
            while idx + len(run) < len(order) and cols_of(int(order[idx + len(run)])).size >= DENSE:
                run.append(int(order[idx + len(run)]))
            if len(run) < 2:
                out.append((i, js)); idx += 1
                continue
            done_before = np.zeros(n, dtype=bool); done_before[np.asarray(order[:idx], dtype=np.int64)] = True
            early = {r: cols_of(r)[done_before[cols_of(r)]] for r in run}
            late = {r: cols_of(r)[~done_before[cols_of(r)]] for r in run}
            nseg = max(-(-early[r].size // SEG) for r in run)
            for k in range(nseg):
                for r in run:
                    seg = early[r][k * SEG:(k + 1) * SEG]
                    if seg.size:
                        out.append((r, seg))
            for r in run:                       # in-run dependencies, in order: each row completes before the next needs it
                out.append((r, late[r]))
            idx += len(run)
        merged = []                               # consecutive visits of one row are one visit
        for i, js in out:
            if merged and merged[-1][0] == i:
                merged[-1] = (i, np.concatenate([merged[-1][1], js]))
            else:
                merged.append((i, js))
        return merged

    def emit_section(order, forward):
        nonlocal e
        visits = visits_of(order, forward)
        sec = [-1 - int(visits[0][0]) * TILE]           # start of the first row (its LU slot is unused)
        e += 1
        last_visit = {}
        for v, (i, js) in enumerate(visits):
            last_visit[i] = v
        for v, (i, js) in enumerate(visits):
            eid[i, js] = e + np.arange(js.size)
            sec.extend((js.astype(np.int64) * TILE).tolist())
            e += js.size
            if not forward:
                diag_eid[i] = e; eid[i, i] = e
            nxt = visits[v + 1][0] if v + 1 < len(visits) else i
            sec.append(-1 - int(nxt) * TILE)
            e += 1
        pad = (-len(sec)) % CH
        sec.extend([SKIP] * pad); e += pad
        sec = np.array(sec, dtype=np.int64)
        return sec
    sec_f = emit_section(order_f, True)
    prog.n_fwd = len(sec_f)
    sec_b = emit_section(order_b, False)
    prog.nE = e
    prog.prog = np.concatenate([sec_f, sec_b]).astype(np.int32)
    prog.ctrl = solve_control_blocks(prog.prog, prog.n_fwd, TILE)
    assert len(prog.prog) == prog.nE
    prog.diag_eid = diag_eid
    # ---- factor program
    fl_ptr = np.zeros(n + 1, dtype=np.int32); fu_ptr = np.zeros(n + 1, dtype=np.int32)
    fl, fu, fd = [], [], []
    for k in range(n):
        li = k + 1 + np.nonzero(F[k + 1:, k])[0]
        uj = k + 1 + np.nonzero(F[k, k + 1:])[0]
        fl.append(eid[li, k]); fu.append(eid[k, uj])
        if li.size and uj.size:
            fd.append(eid[np.ix_(li, uj)].ravel())
        fl_ptr[k + 1] = fl_ptr[k] + li.size
        fu_ptr[k + 1] = fu_ptr[k] + uj.size
    prog.fl_ptr, prog.fu_ptr = fl_ptr, fu_ptr
    prog.fl_eid = np.concatenate(fl).astype(np.int32) if fl else np.zeros(0, np.int32)
    prog.fu_eid = np.concatenate(fu).astype(np.int32) if fu else np.zeros(0, np.int32)
    prog.fd_eid = np.concatenate(fd).astype(np.int32) if fd else np.zeros(0, np.int32)
    assert np.all(prog.fd_eid >= 0)
    # ---- monomials
    monos = sorted(set(batch.cls.templates['M']) | set(batch.cls.templates['L']))
    if not monos:
        monos = [tuple(0 for _ in b.sep_axes)]
    prog.monos = monos
    prog.mono_vals = np.stack([b.monomial_values(batch.cls, m) for m in monos], axis=0)   # (nmono, S)
    midx = {m: i for i, m in enumerate(monos)}
    # ---- template term lists in solver ordering: (row, col, mono, value) for M and L
    def terms(name):
        r, c, m, v = [], [], [], []
        for mono, P in _permuted_templates(batch, name):
            r.append(P.row); c.append(P.col); m.append(np.full(P.nnz, midx[mono])); v.append(P.data)
        if not r:
            return (np.zeros(0, np.int64),) * 3 + (np.zeros(0),)
        return np.concatenate(r), np.concatenate(c), np.concatenate(m), np.concatenate(v)
    prog.terms = {name: terms(name) for name in ('M', 'L')}
    # matvec programs (CSR by row over terms)
    prog.mv = {}
    for name in ('M', 'L'):
        r, c, m, v = prog.terms[name]
        order = np.lexsort((c, r))
        r, c, m, v = r[order], c[order], m[order], v[order]
        ptr = np.zeros(n + 1, dtype=np.int32)
        np.add.at(ptr, r + 1, 1)
        prog.mv[name] = (np.cumsum(ptr).astype(np.int32), c.astype(np.int32), m.astype(np.int32), np.asarray(v, dtype=np.float64))
    batch.eid = eid
    prog.nnz_lu = int(F.sum())
    return prog


def assembly_program(batch, prog, a0, b0):
    """(entry id, mono, value) triples of  a0*M + b0*L  grouped by entry: CSR over entries."""
    es, ms, vs = [], [], []
    for name, w in (('M', a0), ('L', b0)):
        r, c, m, v = prog.terms[name]
        if len(r) == 0:
            continue
        es.append(batch.eid[r, c]); ms.append(m); vs.append(w * v)
    e = np.concatenate(es); m = np.concatenate(ms); v = np.concatenate(vs)
    assert np.all(e >= 0)
    order = np.lexsort((m, e))
    e, m, v = e[order], m[order], v[order]
    ptr = np.zeros(prog.nE + 1, dtype=np.int64)
    np.add.at(ptr, e + 1, 1)
    return np.cumsum(ptr).astype(np.int32), m.astype(np.int32), np.asarray(v, dtype=np.float64)


# ------------------------------------------------------------------------------------------------------
# Gather / scatter maps between field arenas and the SoA pencil vectors
# ------------------------------------------------------------------------------------------------------
class Arena:
    """Concatenated coefficient arrays (local part) of a list of fields / equation outputs."""

    def __init__(self, dist, items):
        """items: list of (tshape, bases)."""
        self.dist = dist
        self.items = items
        self.shapes, self.offsets = [], []
        off = 0
        for tshape, bases in items:
            shp = []
            for ax, bs in enumerate(bases):
                sl = dist.coeff_local_slice(ax, bs)
                shp.append(sl.stop - sl.start)
            self.shapes.append((tuple(tshape), tuple(shp)))
            self.offsets.append(off)
            off += int(np.prod(tshape, dtype=int)) * int(np.prod(shp, dtype=int))
        self.size = off


def line_maps(batch, arena, side):
    """For every coefficient line (slot) of the batch: arena base offset, system-offset kind, length and the
    solver positions of its modes.  side = 'cols' (variables/state) or 'rows' (equations/F).
    Members of a merged batch share everything but the base offset (their parity slot) and the sign of each line:
    line_base / line_sign have shape (R, nlines)."""
    per = [_line_maps_member(batch, arena, side, c) for c in range(batch.R)]
    m = per[0]
    for c, o in enumerate(per[1:], 1):
        same = (np.array_equal(m.line_kind_key, o.line_kind_key) and np.array_equal(m.line_len, o.line_len)
                and np.array_equal(m.line_pos, o.line_pos) and np.array_equal(m.sys_off, o.sys_off))
        if not same:
            raise RuntimeError("merged pencil components do not have aligned coefficient lines")
    m.line_base = np.stack([o.line_base for o in per], axis=0)
    m.line_sign = np.stack([o.line_sign for o in per], axis=0)
    return m


def _line_maps_member(batch, arena, side, member):
    b, cls = batch.builder, batch.cls
    dist = b.dist
    mem = batch.members[member]
    if side == 'cols':
        nat = mem['cols0'][batch.seq] if batch.seq is not None else mem['cols0']
        sign_pos = mem['d2'][batch.seq] if batch.seq is not None else mem['d2']
        slots, offs = cls.col_slots, cls.col_off
    else:
        nat, sign_pos, slots, offs = mem['rows'], mem['d1'], cls.row_slots, cls.row_off
    pos_of_nat = {int(v): i for i, v in enumerate(nat)}
    kinds, kind_tables = {}, []
    lines = []
    g0_start = 0
    if dist.size > 1 and b.sep_axes:
        g0_start = int(b.local_groups(0)[0])
    for q, (slot, o) in enumerate(zip(slots, offs)):
        members = [pos_of_nat.get(int(o + m), -1) for m in range(slot.size)]
        if all(p < 0 for p in members):
            continue
        tshape, bases = arena.items[slot.owner]
        tsh, shp = arena.shapes[slot.owner]
        # strides of the local coefficient array (C order)
        strides = [1] * len(shp)
        for ax in range(len(shp) - 2, -1, -1):
            strides[ax] = strides[ax + 1] * shp[ax + 1]
        comp_stride = int(np.prod(shp, dtype=int))
        base = arena.offsets[slot.owner] + slot.comp * comp_stride
        key = []
        for i, ax in enumerate(b.sep_axes):
            bs = bases[ax]
            if bs is not None:
                base += slot.par[i] * strides[ax]
                key.append((ax, bs.group_size * strides[ax]))
        key = tuple(key)
        if key not in kinds:
            kinds[key] = len(kind_tables)
            off_s = np.zeros(batch.S, dtype=np.int64)
            for ax, st in key:
                i = b.sep_axes.index(ax)
                g = batch.groups[:, i] - (g0_start if ax == 0 else 0)
                off_s += g * st
            kind_tables.append(off_s)
        # a line may belong to this batch only in part (e.g. fully separable problems split per wavenumber
        # block, or invalid modes): emit one entry per maximal run of consecutive member modes
        m = 0
        while m < slot.size:
            if members[m] < 0:
                m += 1
                continue
            m1 = m
            while m1 < slot.size and members[m1] >= 0:
                m1 += 1
            sg = sign_pos[members[m:m1]]
            if np.unique(sg).size != 1:
                raise RuntimeError("sign of a merged pencil component varies along a coefficient line")
            lines.append((base + m, kinds[key], m1 - m, members[m:m1], float(sg[0]), key))
            m = m1
    m = BatchProgram()
    m.line_base = np.array([l[0] for l in lines], dtype=np.int64)
    m.line_sign = np.array([l[4] for l in lines], dtype=np.float64)
    m.line_kind = np.array([l[1] for l in lines], dtype=np.int32)
    m.line_kind_key = np.array([hash(l[5]) for l in lines], dtype=np.int64)
    m.line_len = np.array([l[2] for l in lines], dtype=np.int32)
    m.line_ptr = np.concatenate([[0], np.cumsum(m.line_len)]).astype(np.int32)
    m.line_pos = np.concatenate([np.asarray(l[3], dtype=np.int32) for l in lines]) if lines else np.zeros(0, np.int32)
    m.sys_off = np.stack(kind_tables, axis=0) if kind_tables else np.zeros((1, batch.S), dtype=np.int64)
    return m


def sign_relation(canon, other, rtol=1e-12):
    """If every template of `other` equals D1 T D2 of the canonical component's template (position by position in the
    mode-major orderings, D1 / D2 diagonal +-1), return (d1, d2); else None.  For real-Fourier directions the cos / -sin parity
    blocks of a pencil are related this way: d/dx maps cos -> sin with one sign and sin -> cos with the other
    (reference core/basis.py:1217-1224 group matrix [[0, -k], [k, 0]])."""
    if other.n != canon.n or other.n_border_rows != canon.n_border_rows or other.S != canon.S:
        return None
    for side in ('rows', 'cols'):
        if not np.array_equal(canon.describe(side), other.describe(side)):
            return None
    cls = canon.cls
    n = canon.n
    ii, jj, ss = [], [], []
    for name in ('M', 'L'):
        for mono, T in cls.templates[name].items():
            T = T.tocsr()
            P0 = T[canon.rows][:, canon.cols0].tocsr(); Pc = T[other.rows][:, other.cols0].tocsr()
            P0.sort_indices(); Pc.sort_indices()
            if P0.nnz != Pc.nnz or not np.array_equal(P0.indptr, Pc.indptr) or not np.array_equal(P0.indices, Pc.indices):
                return None
            if P0.nnz == 0:
                continue
            a, b = P0.data, Pc.data
            if np.iscomplexobj(a) or np.iscomplexobj(b):
                return None
            if not np.allclose(np.abs(a), np.abs(b), rtol=rtol, atol=0):
                return None
            coo = P0.tocoo()
            ii.append(coo.row); jj.append(coo.col); ss.append(np.sign(a) * np.sign(b))
    if not ii:
        return None
    ii, jj, ss = np.concatenate(ii), np.concatenate(jj), np.concatenate(ss)
    # breadth-first propagation of  s_ij = d1_i d2_j  over the bipartite pattern (connected: it is one component)
    adj_r = [[] for _ in range(n)]; adj_c = [[] for _ in range(n)]
    for i, j, sg in zip(ii.tolist(), jj.tolist(), ss.tolist()):
        adj_r[i].append((j, sg)); adj_c[j].append((i, sg))
    d1 = np.zeros(n); d2 = np.zeros(n)
    d1[0] = 1.0
    stack = [('r', 0)]
    while stack:
        kind, i = stack.pop()
        if kind == 'r':
            for j, sg in adj_r[i]:
                v = sg * d1[i]
                if d2[j] == 0:
                    d2[j] = v; stack.append(('c', j))
                elif d2[j] != v:
                    return None
        else:
            for r, sg in adj_c[i]:
                v = sg * d2[i]
                if d1[r] == 0:
                    d1[r] = v; stack.append(('r', r))
                elif d1[r] != v:
                    return None
    if np.any(d1 == 0) or np.any(d2 == 0):
        return None
    # the gather / scatter kernels apply ONE sign per coefficient line: the signs must be constant along every line
    for side, d in (('rows', d1), ('cols', d2)):
        desc = canon.describe(side)
        key = desc[:, 0] * 1000003 + desc[:, 1]
        for k in np.unique(key):
            if np.unique(d[key == k]).size != 1:
                return None
    return d1, d2


def merge_sign_equivalent(comps):
    """Greedy grouping of a class's independent components into batches of sign-equivalent members."""
    out = []
    for comp in comps:
        for canon in out:
            rel = sign_relation(canon, comp)
            if rel is not None:
                canon.members.append(dict(rows=comp.rows, cols0=comp.cols0, d1=rel[0], d2=rel[1]))
                break
        else:
            out.append(comp)
    return out


def build_batches(builder, merge=None):
    """One batch per (class, group of sign-equivalent independent components).  merge=False (or DB_MERGE_COMPONENTS=0) keeps
    every component a batch of its own (one factorisation per component, as in round 1)."""
    import os
    if merge is None:
        merge = os.environ.get("DB_MERGE_COMPONENTS", "1") != "0"
    batches = []
    for cls in builder.classes.values():
        if len(cls.groups) == 0 or cls.shape[0] == 0:
            continue
        comps = [Batch(builder, cls, rows, cols) for rows, cols in split_components(cls)]
        batches.extend(merge_sign_equivalent(comps) if merge else comps)
    return batches
