"""Initial value problems on a spherical shell (BASELINE config 5): per-l pencil matrices.

The reference builds one Subproblem per degree l (matrices independent of m; core/basis.py:4187-4192), every operator through
`SphericalEllOperator.subproblem_matrix` (core/operators.py:3159-3193).  Here the LHS expression tree is lowered, for one l at a
time, to blocks  (output component, variable, variable component) -> radial matrix  by restating each operator:
  gradient      xi(-+1, l + R) D-+(l + R)                      core/operators.py:3280-3310
  divergence    xi(-1, l + R + 1) D+ / xi(+1, l + R - 1) D-     core/operators.py:3577-3603
  trace         Q(l)^T (spin trace) Q(l)                        core/operators.py:1783-1827
  products with radial non-constant coefficients (er, rvec)     core/basis.py:284-330 (Gamma through the intertwiners,
                arithmetic.py:560-580) with the radial multiplication matrices of core/basis.py:3879-3915
  lift          e_n (x) Q(l)^T                                   core/basis.py:5155-5199
  interpolation Q(l) (x) row of (dR / r)^k P_n(z(r))             core/basis.py:5823-5889, 3798-3803
  integration   l = 0 row of radial quadrature weights          core/basis.py:5555-5575
  conversions   E^dk, constants -> l = 0 mode                    core/basis.py:4822-4842, 4782-4819
with the radial operators D+-, E of libraries/dedalus_sphere/shell.py:21-73 composed from the Jacobi operators of
dedalus_b200/jacobi.py.  R = regularity total of the component, Q(l) = dedalus_b200/shell.py Intertwiner.
"""
import itertools
import numpy as np
from scipy import sparse
from . import jacobi
from . import operators as ops
from .field import Field
from .shell import ShellBasis, ShellRadialBasis, Intertwiner, shell_basis_of
from .sphere import SphereBasis

REG = (-1, +1, 0)


def xi(mu, l):
    """reference basis.py:3545-3546."""
    return np.sqrt((l + (mu + 1) // 2) / (2 * l + 1))


def regtotal(regindex):
    return sum(REG[i] for i in regindex)


def regularity_allowed(ell, regindex):
    if not regindex:
        return True
    return not Intertwiner(ell).forbidden_regularity(tuple(REG[i] for i in regindex))


class RadialOps:
    """Radial operators of a shell on N modes: D(p, l) and E map basis k -> k + 1 (libraries/dedalus_sphere/shell.py)."""

    def __init__(self, N, radii, alpha=(-0.5, -0.5)):
        self.N, self.radii, self.alpha = N, radii, alpha
        self.dR = radii[1] - radii[0]
        self.rho = (radii[1] + radii[0]) / self.dR
        self._cache = {}

    PAD = 8          # operators are composed on N + PAD modes and truncated ONCE (the reference composes infinite operators)

    def _parts(self, k):
        if k not in self._cache:
            a, b = k + self.alpha[0], k + self.alpha[1]
            P = self.N + self.PAD
            Z = (self.rho * sparse.identity(P) + jacobi.jacobi_matrix(P, a, b)).tocsr()
            AB = jacobi.conversion_matrix(P, a, b, a + 1, b + 1).tocsr()
            Dz = jacobi.differentiation_matrix(P, a, b).tocsr()
            cut = lambda M: sparse.csr_matrix(M.tocsr()[:self.N, :self.N])
            self._cache[k] = dict(DZ=cut(Dz @ Z), AB=cut(AB), E_pad=(0.5 * (AB @ Z)).tocsr(), Z=cut(Z))
        return self._cache[k]

    def D(self, p, l, k):
        """D = (d/dz Z - K AB) / dR with K = (k + 1) + p l + (p == -1) (2 - 3): the reference's K = A(0) - alpha_0 + ... is
        evaluated AFTER AB raised the Jacobi parameter, hence k + 1 (libraries/dedalus_sphere/shell.py:55-67)."""
        q = self._parts(k)
        K = k + 1 + p * l + (-1 if p == -1 else 0)
        return (q['DZ'] - K * q['AB']) / self.dR

    def E(self, k, dk=1):
        """(E^dk)[:N, :N]: E has one sub-diagonal, so the product is formed on padded matrices (dk <= PAD / 2)."""
        P = self.N + self.PAD
        M = sparse.identity(P, format='csr')
        for j in range(dk):
            M = self._parts(k + j)['E_pad'] @ M
        return sparse.csr_matrix(M.tocsr()[:self.N, :self.N])

    def basis_functions(self, k, r):
        """(dR / r)^k P_n^(alpha + k)(z(r)), n < N, at radii r: array (len(r), N)."""
        r = np.atleast_1d(np.asarray(r, dtype=float))
        z = 2 * r / self.dR - self.rho
        return ((self.dR / r) ** k)[:, None] * jacobi.polynomials(self.N, k + self.alpha[0], k + self.alpha[1], z).T

    def multiplication(self, f, k_in, k_out):
        """Matrix of g -> f(r) g from basis k_in to basis k_out by Gauss-Jacobi quadrature (exact for polynomial f times the
        (dR / r)^(k_out - k_in) mismatch being polynomial, which holds for the radial coefficients r and 1 met here)."""
        a, b = k_out + self.alpha[0], k_out + self.alpha[1]
        nq = 2 * self.N + 8
        z, w = jacobi.gauss_grid(nq, a, b)
        r = self.dR / 2 * (z + self.rho)
        Bin = self.basis_functions(k_in, r)                                   # (nq, N)
        Pout = jacobi.polynomials(self.N, a, b, z)                             # (N, nq)
        return (Pout * w[None, :]) @ (((r / self.dR) ** k_out * f(r))[:, None] * Bin)


class Node:
    """Lowered subexpression for one l: blocks[tder][(out comp, var index, var comp)] = radial matrix."""
    __slots__ = ("blocks", "kind", "k", "rank", "nrad")

    def __init__(self, blocks, kind, k, rank, nrad):
        self.blocks, self.kind, self.k, self.rank, self.nrad = blocks, kind, k, rank, nrad


def _map(node, fn, **kw):
    """Apply fn(out comp, matrix) -> list of (new out comp, new matrix) to every block."""
    out = {}
    for t, d in node.blocks.items():
        o = out.setdefault(t, {})
        for (co, iv, ci), B in d.items():
            for co2, B2 in fn(co, B):
                key = (co2, iv, ci)
                o[key] = o[key] + B2 if key in o else B2
    return Node(out, kw.get('kind', node.kind), kw.get('k', node.k), kw.get('rank', node.rank), kw.get('nrad', node.nrad))


class ShellLowering:
    def __init__(self, problem):
        self.problem = problem
        self.variables = problem.variables
        basis = None
        for v in self.variables:
            basis = basis or shell_basis_of(v)
        self.basis = basis
        self.Nr = basis.shape[2]
        self.rops = RadialOps(self.Nr, basis.radii, basis.alpha)
        self._Q = {}

    def Q(self, ell, rank):
        key = (ell, rank)
        if key not in self._Q:
            self._Q[key] = Intertwiner(ell).matrix(rank) if rank else np.ones((1, 1))
        return self._Q[key]

    def kind_of(self, e):
        b = e.bases
        if isinstance(b[-1], ShellBasis):
            return 'shell', b[-1].k
        if isinstance(b[0], (SphereBasis, ShellBasis)):
            return 'sphere', 0
        if isinstance(b[-1], ShellRadialBasis):
            return 'radial', b[-1].k
        return 'const', 0

    # ---- radial NCC: regularity components (at l = 0) of a field on the radial basis, as functions of r
    def ncc_functions(self, ncc):
        """[f_a(r)] over the regularity components a of a radial-basis field given by its coordinate components on the grid."""
        g = np.asarray(ncc['g']).reshape(max(ncc.ncomp, 1), -1)                       # (ncomp, Nr_g) coordinate components
        rank = len(ncc.tensorsig)
        rgrid = ncc.bases[-1].global_grid(ncc.scales[-1])
        if rank:
            cs = ncc.tensorsig[0]
            spin = cs.U_forward(rank) @ g
            if np.abs(spin.imag).max() > 1e-14 * max(np.abs(spin).max(), 1e-300):
                raise NotImplementedError("radial coefficients with angular components")
            reg = self.Q(0, rank).T @ spin.real
        else:
            reg = g
        # interpolating polynomials through the Chebyshev-grid values (the coefficients are low-degree polynomials in r)
        funcs = []
        for a in range(reg.shape[0]):
            coef = np.polynomial.chebyshev.chebfit(2 * rgrid / self.rops.dR - self.rops.rho, reg[a], len(rgrid) - 1)
            coef[np.abs(coef) < 1e-13 * max(np.abs(coef).max(), 1e-300)] = 0
            funcs.append((lambda c: (lambda r: np.polynomial.chebyshev.chebval(2 * r / self.rops.dR - self.rops.rho, c)))(coef)
                         if np.any(coef) else None)
        return funcs

    def gamma(self, ell, rank_a, ell_a, rank_b, ell_b, a_first=True):
        """Tensor-product coupling in regularity components: G[a, b, c] = sum Q_A[a', a] Q_B[b', b] Q_C[(a' b'), c]
        (reference arithmetic.py:560-580 with GammaCoord of the tensor product, 772-783)."""
        QA, QB, QC = self.Q(ell_a, rank_a), self.Q(ell_b, rank_b), self.Q(ell, rank_a + rank_b)
        na, nb = QA.shape[0], QB.shape[0]
        QC3 = QC.reshape(na, nb, -1)                                              # [a', b', c]
        return np.einsum('xa,yb,xyc->abc', QA, QB, QC3)

    # ---- recursive lowering
    def lower(self, e, ell):
        Nr, rops = self.Nr, self.rops
        for iv, v in enumerate(self.variables):
            if e is v:
                kind, k = self.kind_of(v)
                n = Nr if kind == 'shell' else 1
                I = sparse.identity(n, format='csr')
                return Node({0: {(c, iv, c): I for c in range(max(v.ncomp, 1))}}, kind, k, len(v.tensorsig), n)
        if isinstance(e, ops.Add):
            nodes = [self.lower(a, ell) for a in e.args]
            kind = 'shell' if any(n.kind == 'shell' for n in nodes) else nodes[0].kind
            k = max(n.k for n in nodes)
            out = {}
            for n in nodes:
                n = self.convert(n, kind, k, ell)
                for t, d in n.blocks.items():
                    o = out.setdefault(t, {})
                    for key, B in d.items():
                        o[key] = o[key] + B if key in o else B
            return Node(out, kind, k, nodes[0].rank, Nr if kind == 'shell' else 1)
        if isinstance(e, ops.ScalarMul):
            return _map(self.lower(e.args[0], ell), lambda co, B: [(co, B * e.c)])
        if isinstance(e, ops.TimeDerivative):
            n = self.lower(e.args[0], ell)
            return Node({t + 1: d for t, d in n.blocks.items()}, n.kind, n.k, n.rank, n.nrad)
        if isinstance(e, ops.Gradient):
            n = self.lower(e.args[0], ell)
            if n.kind != 'shell':
                raise NotImplementedError("gradient of a non-shell operand")
            ncomp = 3 ** n.rank
            def fn(co, B):
                rin = np.unravel_index(co, (3,) * n.rank) if n.rank else ()
                R = regtotal(rin)
                out = []
                for a, mu in ((0, -1), (1, +1)):
                    rout = (a,) + tuple(rin)
                    if regularity_allowed(ell, rin) and regularity_allowed(ell, rout):
                        out.append((a * ncomp + co, xi(mu, ell + R) * (rops.D(mu, ell + R, n.k) @ B)))
                return out
            return _map(n, fn, k=n.k + 1, rank=n.rank + 1)
        if isinstance(e, ops.Divergence):
            n = self.lower(e.args[0], ell)
            nrest = 3 ** (n.rank - 1)
            def fn(co, B):
                rin = tuple(np.unravel_index(co, (3,) * n.rank))
                R = regtotal(rin)
                a, rest = rin[0], rin[1:]
                if a == 2 or not (regularity_allowed(ell, rin) and regularity_allowed(ell, rest)):
                    return []
                if a == 0:
                    M = xi(-1, ell + R + 1) * rops.D(+1, ell + R, n.k)
                else:
                    M = xi(+1, ell + R - 1) * rops.D(-1, ell + R, n.k)
                return [(co % nrest, M @ B)]
            return _map(n, fn, k=n.k + 1, rank=n.rank - 1)
        if isinstance(e, ops.Trace):
            n = self.lower(e.args[0], ell)
            nrest = 3 ** (n.rank - 2)
            tr = np.zeros(9); tr[[1, 3, 8]] = 1                                  # spin trace: (-+), (+-), (00)
            T = self.Q(ell, n.rank - 2).T @ np.kron(tr[None, :], np.eye(nrest)) @ self.Q(ell, n.rank)
            def fn(co, B):
                return [(c, T[c, co] * B) for c in range(nrest) if T[c, co] != 0]
            return _map(n, fn, rank=n.rank - 2)
        if isinstance(e, ops.Multiply):
            A, Bx = e.args
            a_ncc = not A.has(*self.variables)
            ncc, arg = (A, Bx) if a_ncc else (Bx, A)
            if ncc.has(*self.variables):
                raise ValueError("LHS must be linear in the problem variables.")
            if not isinstance(ncc, Field):
                raise NotImplementedError("only plain fields as LHS coefficients")
            n = self.lower(arg, ell)
            kind_n, k_ncc = self.kind_of(ncc)
            if kind_n == 'const':
                vals = np.asarray(ncc['c']).reshape(-1)
                if len(vals) != 1:
                    raise NotImplementedError("constant tensor coefficients on a shell")
                return _map(n, lambda co, B: [(co, B * vals[0])])
            if kind_n != 'radial' or n.kind != 'shell':
                raise NotImplementedError("LHS coefficients must be constants or fields on the shell's radial basis")
            funcs = self.ncc_functions(ncc)
            ra, rb = len(ncc.tensorsig), n.rank
            if a_ncc:
                G = self.gamma(ell, ra, 0, rb, ell)                             # [ncc comp, arg comp, out comp]
            else:
                G = self.gamma(ell, rb, ell, ra, 0).transpose(1, 0, 2)
            k_out = n.k + k_ncc
            mats = [None if f is None else sparse.csr_matrix(_clean(rops.multiplication(f, n.k, k_out))) for f in funcs]
            def fn(co, B):
                out = []
                for c in range(G.shape[2]):
                    M = None
                    for a, Ma in enumerate(mats):
                        if Ma is not None and abs(G[a, co, c]) > 1e-14:
                            M = G[a, co, c] * Ma if M is None else M + G[a, co, c] * Ma
                    if M is not None:
                        out.append((c, M @ B))
                return out
            return _map(n, fn, k=k_out, rank=ra + rb)
        if isinstance(e, ops.Lift):
            n = self.lower(e.args[0], ell)
            if n.kind != 'sphere' or not isinstance(e.basis, ShellBasis):
                raise NotImplementedError("Lift of a sphere field along a shell basis only")
            pos = e.n if e.n >= 0 else e.n + Nr
            col = sparse.csr_matrix(([1.0], ([pos], [0])), shape=(Nr, 1))
            Q = self.Q(ell, n.rank)
            def fn(co, B):
                return [(c, Q[co, c] * (col @ B)) for c in range(Q.shape[1]) if Q[co, c] != 0]
            return _map(n, fn, kind='shell', k=e.basis.k, nrad=Nr)
        if isinstance(e, ops.Interpolate):
            n = self.lower(e.args[0], ell)
            if n.kind != 'shell' or e.axis != e.dist.dim - 1:
                raise NotImplementedError("interpolation along the radius of a shell only")
            row = sparse.csr_matrix(rops.basis_functions(n.k, [float(e.position)]))   # (1, Nr)
            Q = self.Q(ell, n.rank)
            def fn(co, B):
                rin = tuple(np.unravel_index(co, (3,) * n.rank)) if n.rank else ()
                if not regularity_allowed(ell, rin):
                    return []
                return [(s, Q[s, co] * (row @ B)) for s in range(Q.shape[0]) if Q[s, co] != 0]
            return _map(n, fn, kind='sphere', k=0, nrad=1)
        if isinstance(e, ops.Integrate):
            n = self.lower(e.args[0], ell)
            if n.kind != 'shell' or n.rank != 0 or e.average:
                raise NotImplementedError("integration of shell scalars only")
            if ell != 0:
                return Node({}, 'const', 0, 0, 1)
            z, w = np.polynomial.legendre.leggauss(2 * Nr)
            r = rops.dR / 2 * (z + rops.rho)
            Qk = jacobi.polynomials(Nr, n.k + rops.alpha[0], n.k + rops.alpha[1], z)
            row = ((r**2 * w * (r / rops.dR) ** (-n.k))[None, :] @ Qk.T) * (rops.dR / 2) * (4 * np.pi / np.sqrt(2))
            row = sparse.csr_matrix(row)
            return _map(n, lambda co, B: [(co, row @ B)], kind='const', k=0, nrad=1)
        raise NotImplementedError(f"{type(e).__name__} is not supported on the LHS of shell problems")

    def convert(self, n, kind, k, ell):
        Nr, rops = self.Nr, self.rops
        if n.kind == kind and n.k == k:
            return n
        if n.kind == 'shell' and kind == 'shell':
            E = rops.E(n.k, k - n.k)
            return _map(n, lambda co, B: [(co, E @ B)], k=k)
        if n.kind == 'const' and kind == 'shell':
            if n.rank:
                raise NotImplementedError("constant tensors on a shell")
            if ell != 0:
                return Node({}, 'shell', k, 0, Nr)
            cmv = (1 / np.sqrt(jacobi.mass(rops.alpha[0], rops.alpha[1]))) / np.sqrt(2)      # reference basis.py:4203-4207
            col = sparse.csr_matrix(([1 / cmv], ([0], [0])), shape=(Nr, 1))
            E = rops.E(0, k)
            return _map(n, lambda co, B: [(co, E @ col @ B)], kind='shell', k=k, nrad=Nr)
        raise NotImplementedError(f"conversion {n.kind} -> {kind}")


def _clean(M, tol=1e-13):
    M = np.array(M)
    M[np.abs(M) < tol * max(np.abs(M).max(), 1e-300)] = 0
    return M


def component_layout(items, basis, ell):
    """Natural ordering of the unknowns (or equation rows) of degree l: for each item (kind, rank), its components, each with its
    radial size, and their validity (reference: regularity_allowed / |s| <= l / l == 0)."""
    Nr = basis.shape[2]
    out = []
    for it, (kind, rank) in enumerate(items):
        ncomp = 3 ** rank
        for c in range(ncomp):
            idx = tuple(np.unravel_index(c, (3,) * rank)) if rank else ()
            if kind == 'shell':
                valid, n = regularity_allowed(ell, idx), Nr
            elif kind == 'sphere':
                valid, n = abs(sum(REG[i] for i in idx)) <= ell, 1
            else:
                valid, n = (ell == 0), 1
            out.append(dict(item=it, comp=c, n=n, valid=valid))
    return out


def assemble(low, ell):
    """(M, L, row layout, column layout) of degree l in natural ordering: dense arrays over ALL components (invalid ones
    included, as zero rows / columns)."""
    problem = low.problem
    var_items = [(low.kind_of(v)[0], len(v.tensorsig)) for v in low.variables]
    cols = component_layout(var_items, low.basis, ell)
    col_off = np.cumsum([0] + [c['n'] for c in cols])
    col_index = {(c['item'], c['comp']): i for i, c in enumerate(cols)}
    eq_nodes = [low.lower(eq['LHS'], ell) for eq in problem.equations]
    eq_items = [(low.kind_of(eq['LHS'])[0], len(eq['tensorsig'])) for eq in problem.equations]
    rows = component_layout(eq_items, low.basis, ell)
    row_off = np.cumsum([0] + [r['n'] for r in rows])
    row_index = {(r['item'], r['comp']): i for i, r in enumerate(rows)}
    mats = {0: np.zeros((row_off[-1], col_off[-1])), 1: np.zeros((row_off[-1], col_off[-1]))}
    for ie, node in enumerate(eq_nodes):
        for t, d in node.blocks.items():
            if t > 1:
                raise NotImplementedError("Only first-order time derivatives are supported.")
            for (co, iv, ci), B in d.items():
                ri, cj = row_index[(ie, co)], col_index[(iv, ci)]
                B = B.toarray() if sparse.issparse(B) else np.asarray(B)
                mats[t][row_off[ri]:row_off[ri + 1], col_off[cj]:col_off[cj + 1]] += B
    return mats[1], mats[0], rows, cols
