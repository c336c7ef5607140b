"""CPU restatement of the reference's Rayleigh-Benard IVP step (2-D: examples/ivp_2d_rayleigh_benard/
rayleigh_benard.py:33-89; 3-D: the same equations with a y axis).  TEST INFRASTRUCTURE.

It follows the reference's flow, not the product's:
  * one sparse matrix pair (M, L) per (kx[,ky]) pencil in the reference's natural ordering
    [p, b, u, tau_p, tau_b1, tau_b2, tau_u1, tau_u2] x (component, parity_x[, parity_y], n)
    (core/subsystems.py:497-602; block contents per SURVEY.md Appendix B), invalid sin(0) modes dropped;
  * per-pencil SuperLU factorisation of the transposed LHS and solve (libraries/matsolvers.py:126-183);
  * the IMEX stage loops of core/timesteppers.py:95-187 and 552-644;
  * the explicit term F = (-u.grad b, -u.grad u) evaluated by coefficient-space derivatives, backward
    transforms at dealias 3/2, grid products, forward transforms (core/evaluator.py:95-146), through either the
    matrix transforms or the scipy-FFT transforms of oracle/transforms_oracle.py.
"""
import itertools
import numpy as np
from scipy import sparse
from scipy.sparse.linalg import splu
from . import transforms_oracle as T
from . import imex

J2 = np.array([[0., -1.], [1., 0.]])
I2 = np.eye(2)


class RBOracle:
    def __init__(self, dim, Nh, Nz, Ra, Pr=1.0, Lx=4.0, Lz=1.0, dealias=1.5, transforms="fft"):
        self.dim, self.Nh, self.Nz = dim, Nh, Nz
        self.h = dim - 1
        self.Lx, self.Lz = Lx, Lz
        self.kappa = (Ra * Pr) ** -0.5
        self.nu = (Ra / Pr) ** -0.5
        self.Gh = int(dealias * Nh); self.Gz = int(dealias * Nz)
        self.transforms = transforms
        self.kfund = 2 * np.pi / Lx
        st = Lz / 2.0                                     # affine stretch of the Chebyshev interval
        self.D0 = sparse.csr_matrix(T.jacobi_differentiation(Nz, -0.5, -0.5)) / st
        self.D1 = sparse.csr_matrix(T.jacobi_differentiation(Nz, 0.5, 0.5)) / st
        self.C0 = sparse.csr_matrix(self._clean(T.jacobi_conversion(Nz, -0.5, -0.5, 0.5, 0.5)))
        self.C1 = sparse.csr_matrix(self._clean(T.jacobi_conversion(Nz, 0.5, 0.5, 1.5, 1.5)))
        self.D0 = sparse.csr_matrix(self._clean(self.D0.toarray()))
        self.D1 = sparse.csr_matrix(self._clean(self.D1.toarray()))
        self.C02 = (self.C1 @ self.C0).tocsr()
        self.i_left = T.jacobi_polynomials(Nz, -0.5, -0.5, np.array([-1.0]))[:, 0][None, :]
        self.i_right = T.jacobi_polynomials(Nz, -0.5, -0.5, np.array([1.0]))[:, 0][None, :]
        self.integ = (T.jacobi_integration(Nz, -0.5, -0.5) * st)[None, :]
        self.lift = sparse.csr_matrix(([1.0], ([Nz - 1], [0])), shape=(Nz, 1))
        self._build_transform_matrices()
        self.pencils = {}
        self.lu_cache = {}

    @staticmethod
    def _mass(a, b):
        from scipy.special import gammaln
        return np.exp((a + b + 1) * np.log(2.0) + gammaln(a + 1) + gammaln(b + 1) - gammaln(a + b + 2))

    @staticmethod
    def _clean(A):
        A = np.array(A)
        A[np.abs(A) < 1e-13 * np.abs(A).max()] = 0
        return A

    # ---- field shapes --------------------------------------------------------------------------------------
    @property
    def cshape(self):
        return (self.Nh,) * self.h + (self.Nz,)

    def groups(self):
        return list(itertools.product(range(self.Nh // 2), repeat=self.h))

    # ---- pencil matrices -----------------------------------------------------------------------------------
    def _hor(self, mats):
        out = np.array([[1.0]])
        for m in mats:
            out = np.kron(out, m)
        return sparse.csr_matrix(out)

    def pencil_matrices(self, grp):
        """(M, L, valid mask) in natural ordering with ALL parity slots (invalid rows/cols zero, like
        pre_left.T @ X_min @ pre_right.T of the reference)."""
        dim, h, Nz = self.dim, self.h, self.Nz
        k = [g * self.kfund for g in grp]
        P = 2 ** h                                   # parity slots
        Ih = self._hor([I2] * h)
        G = [self._hor([k[a] * J2 if a == b else I2 for b in range(h)]) for a in range(h)]
        lap_h = sum((Ga @ Ga for Ga in G), sparse.csr_matrix((P, P)))
        zero0 = all(g == 0 for g in grp)
        C0, C1, C02, D0, D1, lift = self.C0, self.C1, self.C02, self.D0, self.D1, self.lift
        kr = lambda A, B: sparse.kron(A, B, format='csr')
        nb = P * Nz
        # variable blocks: name -> (offset, size)
        sizes = [('p', nb), ('b', nb)] + [(f'u{j}', nb) for j in range(dim)] + [('tau_p', 1 if zero0 else 0), ('tau_b1', P), ('tau_b2', P)]
        sizes += [(f'tau_u1{j}', P) for j in range(dim)] + [(f'tau_u2{j}', P) for j in range(dim)]
        voff, o = {}, 0
        for nme, s in sizes:
            voff[nme] = (o, s); o += s
        ncol = o
        esizes = [('cont', nb), ('beq', nb)] + [(f'ueq{j}', nb) for j in range(dim)] + [('b_l', P)] + [(f'u_l{j}', P) for j in range(dim)]
        esizes += [('b_r', P)] + [(f'u_r{j}', P) for j in range(dim)] + [('gauge', 1 if zero0 else 0)]
        eoff, o = {}, 0
        for nme, s in esizes:
            eoff[nme] = (o, s); o += s
        nrow = o
        Mb, Lb = {}, {}

        def add(dct, e, v, blk):
            dct[(e, v)] = dct[(e, v)] + blk if (e, v) in dct else blk
        LapT = kr(lap_h, C02) + kr(Ih, D1 @ D0)                       # div(grad .) of a T-basis scalar, in (3/2,3/2)
        lift32 = C1 @ lift                                             # lift(tau) converted to (3/2,3/2)
        dlift32 = D1 @ lift                                            # dz(lift(tau)) in (3/2,3/2)
        # continuity: trace(grad_u) + tau_p = 0        [(1/2,1/2) basis]
        for a in range(h):
            add(Lb, 'cont', f'u{a}', kr(G[a], C0))
        add(Lb, 'cont', f'u{dim - 1}', kr(Ih, D0))
        add(Lb, 'cont', f'tau_u1{dim - 1}', kr(Ih, lift))
        if zero0:
            col = np.zeros((nb, 1)); col[0, 0] = np.sqrt(self._mass(0.5, 0.5))
            add(Lb, 'cont', 'tau_p', sparse.csr_matrix(col))
        # buoyancy
        add(Mb, 'beq', 'b', kr(Ih, C02))
        add(Lb, 'beq', 'b', -self.kappa * LapT)
        add(Lb, 'beq', 'tau_b1', -self.kappa * kr(Ih, dlift32))
        add(Lb, 'beq', 'tau_b2', kr(Ih, lift32))
        # momentum
        for j in range(dim):
            add(Mb, f'ueq{j}', f'u{j}', kr(Ih, C02))
            add(Lb, f'ueq{j}', f'u{j}', -self.nu * LapT)
            add(Lb, f'ueq{j}', f'tau_u1{j}', -self.nu * kr(Ih, dlift32))
            add(Lb, f'ueq{j}', f'tau_u2{j}', kr(Ih, lift32))
            if j < h:
                add(Lb, f'ueq{j}', 'p', kr(G[j], C02))
            else:
                add(Lb, f'ueq{j}', 'p', kr(Ih, C1 @ D0))
                add(Lb, f'ueq{j}', 'b', -kr(Ih, C02))
        # boundary conditions
        add(Lb, 'b_l', 'b', kr(Ih, sparse.csr_matrix(self.i_left)))
        add(Lb, 'b_r', 'b', kr(Ih, sparse.csr_matrix(self.i_right)))
        for j in range(dim):
            add(Lb, f'u_l{j}', f'u{j}', kr(Ih, sparse.csr_matrix(self.i_left)))
            add(Lb, f'u_r{j}', f'u{j}', kr(Ih, sparse.csr_matrix(self.i_right)))
        if zero0:
            row = np.zeros((1, nb)); row[0, :Nz] = self.integ[0] * (self.Lx ** h)
            add(Lb, 'gauge', 'p', sparse.csr_matrix(row))

        def assemble(blocks):
            A = sparse.lil_matrix((nrow, ncol))
            for (e, v), blk in blocks.items():
                (r0, rs), (c0, cs) = eoff[e], voff[v]
                if rs and cs:
                    A[r0:r0 + rs, c0:c0 + cs] = blk
            return A.tocsr()
        M, L = assemble(Mb), assemble(Lb)
        # valid modes: drop sin part of any k=0 axis (core/basis.py:1123-1134)
        par = list(itertools.product(range(2), repeat=h))
        pvalid = np.array([all(not (grp[a] == 0 and pp[a] == 1) for a in range(h)) for pp in par])
        def mask(sz_list):
            m = []
            for nme, s in sz_list:
                if s == nb:
                    m.append(np.repeat(pvalid, Nz))
                elif s == P:
                    m.append(pvalid)
                else:
                    m.append(np.ones(s, dtype=bool))
            return np.concatenate(m)
        vr, vc = mask(esizes), mask(sizes)
        M = sparse.diags(vr.astype(float)) @ M @ sparse.diags(vc.astype(float))
        L = sparse.diags(vr.astype(float)) @ L @ sparse.diags(vc.astype(float))
        for A in (M, L):
            A.data[np.abs(A.data) < 1e-12] = 0
            A.eliminate_zeros()
        return M.tocsr(), L.tocsr(), vr, vc, voff, eoff

    # ---- transforms ----------------------------------------------------------------------------------------
    def _build_transform_matrices(self):
        if self.transforms == "matrix":
            self.rf_f, self.rf_b = T.rf_matrices(self.Gh, self.Nh)
            self.ch_b = {al: T.jacobi_matrices(self.Gz, self.Nz, al - 0.5, al - 0.5, -0.5, -0.5)[1] for al in (0, 1)}
            self.ch_f = {al: T.jacobi_matrices(self.Gz, self.Nz, al - 0.5, al - 0.5, -0.5, -0.5)[0] for al in (0, 2)}

    def to_grid(self, c, alpha=0):
        """coefficients (Nh.., Nz) in basis alpha -> dealiased grid (axes last -> first, distributor.py:131-175)."""
        z = self.h
        if self.transforms == "matrix":
            g = T.apply_along(self.ch_b[alpha], c, c.ndim - 1)
            for ax in range(c.ndim - 2, c.ndim - 2 - self.h, -1):
                g = T.apply_along(self.rf_b, g, ax)
            return g
        g = T.cheb_backward_fft(c, self.Gz, c.ndim - 1, alpha - 0.5, alpha - 0.5)
        for ax in range(c.ndim - 2, c.ndim - 2 - self.h, -1):
            g = T.rf_backward_fft(g, self.Gh, ax)
        return g

    def to_coeff(self, g, alpha=2):
        if self.transforms == "matrix":
            c = g
            for ax in range(g.ndim - 1 - self.h, g.ndim - 1):
                c = T.apply_along(self.rf_f, c, ax)
            return T.apply_along(self.ch_f[alpha], c, c.ndim - 1)
        c = g
        for ax in range(g.ndim - 1 - self.h, g.ndim - 1):
            c = T.rf_forward_fft(c, self.Nh, ax)
        return T.cheb_forward_fft(c, self.Nz, c.ndim - 1, alpha - 0.5, alpha - 0.5)

    def hderiv(self, c, axis):
        """d/dx in coefficient space: group matrix [[0,-k],[k,0]] (core/basis.py:1203-1224)."""
        out = np.zeros_like(c)
        k = np.arange(self.Nh // 2) * self.kfund
        shape = [1] * c.ndim; shape[axis] = -1
        kk = k.reshape(shape)
        cos = c[T.axslice(axis, 0, None, 2)]; msin = c[T.axslice(axis, 1, None, 2)]
        out[T.axslice(axis, 0, None, 2)] = -kk * msin
        out[T.axslice(axis, 1, None, 2)] = kk * cos
        return out

    def rhs(self, b, u):
        """F_b = -u.grad(b), F_u = -u.grad(u) in the (3/2,3/2) equation basis; b (..Nz), u (dim, .., Nz) in T."""
        dim, h = self.dim, self.h
        zax = b.ndim - 1
        ug = [self.to_grid(u[j], 0) for j in range(dim)]
        def grad_grid(f):
            comps = []
            for a in range(h):
                # horizontal derivative converted to the derivative basis (1/2,1/2) like CartesianGradient
                comps.append(self.to_grid(T.apply_along(self.C0.toarray(), self.hderiv(f, a), zax), 1))
            comps.append(self.to_grid(T.apply_along(self.D0.toarray(), f, zax), 1))
            return comps
        gb = grad_grid(b)
        Fb = self.to_coeff(-sum(ug[i] * gb[i] for i in range(dim)), 2)
        Fu = []
        for j in range(dim):
            gu = grad_grid(u[j])
            Fu.append(self.to_coeff(-sum(ug[i] * gu[i] for i in range(dim)), 2))
        return Fb, np.stack(Fu)

    # ---- pencil gather / scatter ---------------------------------------------------------------------------
    def _slices(self, grp):
        return tuple(slice(2 * g, 2 * g + 2) for g in grp)

    def gather_state(self, grp, st, voff):
        P = 2 ** self.h; Nz = self.Nz
        x = np.zeros(sum(s for _, s in voff.values()))
        sl = self._slices(grp)
        def put(name, arr):
            o, s = voff[name]
            if s:
                x[o:o + s] = arr.reshape(-1)[:s]
        put('p', st['p'][sl]); put('b', st['b'][sl])
        for j in range(self.dim):
            put(f'u{j}', st['u'][j][sl])
            put(f'tau_u1{j}', st['tau_u1'][j][sl]); put(f'tau_u2{j}', st['tau_u2'][j][sl])
        put('tau_b1', st['tau_b1'][sl]); put('tau_b2', st['tau_b2'][sl])
        if voff['tau_p'][1]:
            x[voff['tau_p'][0]] = st['tau_p']
        return x

    def scatter_state(self, grp, x, st, voff):
        sl = self._slices(grp)
        def get(name, shape):
            o, s = voff[name]
            return x[o:o + s].reshape(shape)
        P2 = (2,) * self.h
        st['p'][sl] = get('p', P2 + (self.Nz,)); st['b'][sl] = get('b', P2 + (self.Nz,))
        for j in range(self.dim):
            st['u'][j][sl] = get(f'u{j}', P2 + (self.Nz,))
            st['tau_u1'][j][sl] = get(f'tau_u1{j}', P2); st['tau_u2'][j][sl] = get(f'tau_u2{j}', P2)
        st['tau_b1'][sl] = get('tau_b1', P2); st['tau_b2'][sl] = get('tau_b2', P2)
        if voff['tau_p'][1]:
            st['tau_p'] = x[voff['tau_p'][0]]

    def gather_F(self, grp, Fb, Fu, eoff):
        f = np.zeros(sum(s for _, s in eoff.values()))
        sl = self._slices(grp)
        o, s = eoff['beq']; f[o:o + s] = Fb[sl].reshape(-1)
        for j in range(self.dim):
            o, s = eoff[f'ueq{j}']; f[o:o + s] = Fu[j][sl].reshape(-1)
        if all(g == 0 for g in grp):
            o, s = eoff['b_l']; f[o] = self.Lz          # b(z=0) = Lz : constant -> cos(0)cos(0) slot
        return f

    # ---- time stepping -------------------------------------------------------------------------------------
    def new_state(self, b0_c):
        hs = (self.Nh,) * self.h
        return dict(p=np.zeros(hs + (self.Nz,)), b=np.array(b0_c, dtype=float, copy=True),
                    u=np.zeros((self.dim,) + hs + (self.Nz,)), tau_p=0.0,
                    tau_b1=np.zeros(hs), tau_b2=np.zeros(hs),
                    tau_u1=np.zeros((self.dim,) + hs), tau_u2=np.zeros((self.dim,) + hs))

    def _pencil(self, grp):
        if grp not in self.pencils:
            self.pencils[grp] = self.pencil_matrices(grp)
        return self.pencils[grp]

    def _solve(self, grp, key, a0, b0, rhs):
        M, L, vr, vc, voff, eoff = self._pencil(grp)
        ck = (grp, key)
        if ck not in self.lu_cache:
            A = (a0 * M + b0 * L)[vr][:, vc].tocsc()
            self.lu_cache[ck] = splu(A.T.tocsc())            # SuperluColamdFactorizedTranspose, matsolvers.py:179-183
        x = np.zeros(len(vc))
        x[vc] = self.lu_cache[ck].solve(rhs[vr], trans='T')
        return x

    def step_rk(self, st, dt, scheme="RK222", groups=None):
        tab = imex.RK[scheme]; A, H = tab['A'], tab['H']
        stages = len(tab['c']) - 1
        groups = self.groups() if groups is None else groups
        MX0, LX, Fs = {}, [dict() for _ in range(stages)], [dict() for _ in range(stages)]
        for grp in groups:
            M, L, vr, vc, voff, eoff = self._pencil(grp)
            x = self.gather_state(grp, st, voff)
            MX0[grp] = M @ x; LX[0][grp] = L @ x
        for i in range(1, stages + 1):
            if i > 1:
                for grp in groups:
                    M, L, vr, vc, voff, eoff = self._pencil(grp)
                    LX[i - 1][grp] = L @ self.gather_state(grp, st, voff)
            Fb, Fu = self.rhs(st['b'], st['u'])
            for grp in groups:
                M, L, vr, vc, voff, eoff = self._pencil(grp)
                Fs[i - 1][grp] = self.gather_F(grp, Fb, Fu, eoff)
                rhs = MX0[grp].copy()
                for j in range(i):
                    rhs += dt * A[i, j] * Fs[j][grp] - dt * H[i, j] * LX[j][grp]
                x = self._solve(grp, (dt, H[i, i]), 1.0, dt * H[i, i], rhs)
                self.scatter_state(grp, x, st, voff)
        return st

    def step_sbdf2(self, st, dt, hist):
        """hist: dict with 'it', 'dts', 'MX', 'LX', 'F' lists (most recent first)."""
        groups = self.groups()
        hist['dts'] = [dt] + hist['dts'][:1]
        a, b, c = imex.sbdf2(hist['dts'][0], hist['dts'][1] if len(hist['dts']) > 1 else dt, hist['it'])
        hist['it'] += 1
        MX, LX, F = {}, {}, {}
        for grp in groups:
            M, L, vr, vc, voff, eoff = self._pencil(grp)
            x = self.gather_state(grp, st, voff)
            MX[grp] = M @ x; LX[grp] = L @ x
        Fb, Fu = self.rhs(st['b'], st['u'])
        for grp in groups:
            M, L, vr, vc, voff, eoff = self._pencil(grp)
            F[grp] = self.gather_F(grp, Fb, Fu, eoff)
        hist['MX'] = [MX] + hist['MX'][:1]; hist['LX'] = [LX] + hist['LX'][:1]; hist['F'] = [F] + hist['F'][:1]
        for grp in groups:
            M, L, vr, vc, voff, eoff = self._pencil(grp)
            rhs = np.zeros(M.shape[0])
            for j in range(1, len(c)):
                if c[j] != 0: rhs += c[j] * hist['F'][j - 1][grp]
            for j in range(1, len(a)):
                if a[j] != 0: rhs -= a[j] * hist['MX'][j - 1][grp]
            for j in range(1, len(b)):
                if b[j] != 0: rhs -= b[j] * hist['LX'][j - 1][grp]
            x = self._solve(grp, (a[0], b[0]), a[0], b[0], rhs)
            self.scatter_state(grp, x, st, voff)
        return st


def run(dim, Nh, Nz, Ra, b0_c, steps, dt, scheme="RK222", transforms="fft", u0_c=None):
    """`dt` may be a sequence (one value per step: every change drops the factorisations like the reference,
    core/timesteppers.py:577-583); `u0_c` an initial velocity (default 0, the stock scripts' initial condition)."""
    orc = RBOracle(dim, Nh, Nz, Ra, transforms=transforms)
    st = orc.new_state(b0_c)
    if u0_c is not None:
        st['u'] = np.array(u0_c, dtype=float, copy=True)
    hist = dict(it=0, dts=[], MX=[], LX=[], F=[])
    dts = list(dt) if np.ndim(dt) else [dt] * steps
    for i in range(steps):
        if i and dts[i] != dts[i - 1]:
            orc.lu_cache.clear()
        if scheme == "SBDF2":
            orc.step_sbdf2(st, dts[i], hist)
        else:
            orc.step_rk(st, dts[i], scheme)
    return st
