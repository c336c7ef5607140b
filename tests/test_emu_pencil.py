"""csrc/pencil.cu + pointwise.cu logic under the test-only CPU emulation vs the numpy program interpreters /
direct numpy evaluation."""
import numpy as np, pytest, ctypes as C
from scipy import sparse
from scipy.sparse.linalg import spsolve
from dedalus_b200 import examples
from dedalus_b200 import lib as dlib
from dedalus_b200.pencils import (PencilSystemBuilder, build_batches, compile_batch, assembly_program, Arena, line_maps)
from emu import emu_lib as E
import program_interp as pi

GAMMA = (2 - np.sqrt(2)) / 2


def _i32(a): return np.ascontiguousarray(a, dtype=np.int32)


def test_pencil_kernels_match_interpreters():
    lib = E.emu()
    pb = examples.rayleigh_benard(dim=3, Nh=6, Nz=10)
    builder = PencilSystemBuilder(pb['problem'])
    batches = build_batches(builder, merge=False)      # single-component batches: the per-batch kernels and interpreters
    a0, b0 = 1.0, 0.02 * GAMMA
    rng = np.random.default_rng(1)
    var_arena = Arena(builder.dist, [(v.tshape, v.bases) for v in builder.variables])
    state = rng.standard_normal(var_arena.size)
    state_out = np.zeros_like(state)
    covered = np.zeros(var_arena.size, dtype=bool)
    for batch in batches:
        prog = compile_batch(batch, a0, b0)
        n, S = prog.n, prog.S
        ld = prog.ld
        mono = np.zeros((len(prog.monos), ld)); mono[:, :S] = prog.mono_vals
        # --- assemble + factor
        aptr, amono, aval = assembly_program(batch, prog, a0, b0)
        LUt = np.zeros(prog.nE * ld)
        lib.call("db_pencil_assemble", E.ptr(LUt), prog.nE, S, ld, E.ptr(mono), E.ptr(_i32(aptr)), E.ptr(_i32(amono)), E.ptr(aval), None)
        ref_LU = pi.assemble(prog, (aptr, amono, aval))
        assert np.allclose(pi.from_tiles(LUt, prog.nE, ld)[:, :S], ref_LU, rtol=0, atol=0)
        info = np.zeros(1, dtype=np.int32)
        lib.call("db_pencil_factor", E.ptr(LUt), n, S, prog.nE, E.ptr(_i32(prog.diag_eid)), E.ptr(_i32(prog.fl_ptr)), E.ptr(_i32(prog.fl_eid)),
                 E.ptr(_i32(prog.fu_ptr)), E.ptr(_i32(prog.fu_eid)), E.ptr(_i32(prog.fd_eid)), E.ptr(info), None)
        assert info[0] == 0
        ref_LU = pi.factor(prog, ref_LU)
        assert np.allclose(pi.from_tiles(LUt, prog.nE, ld)[:, :S], ref_LU, rtol=1e-13, atol=1e-13)
        # --- gather state -> X, matvec, solve with a fused 3-term RHS, scatter back
        maps = line_maps(batch, var_arena, 'cols')
        Xt = np.zeros(n * ld)
        kind_ext = _i32(maps.line_kind)
        sys_off = np.zeros((maps.sys_off.shape[0], ld), dtype=np.int64); sys_off[:, :S] = maps.sys_off
        max_len = int(maps.line_len.max())
        lib.call("db_pencil_gather", E.ptr(state), E.ptr(Xt), S, n, maps.line_base.shape[1], max_len, E.ptr(np.ascontiguousarray(maps.line_base[0])), E.ptr(kind_ext),
                 E.ptr(_i32(maps.line_ptr)), E.ptr(_i32(maps.line_pos)), E.ptr(sys_off), ld, None)
        Xref = pi.gather(maps, state, n, S)
        X = pi.from_tiles(Xt, n, ld)
        assert np.array_equal(X[:, :S], Xref)
        yMt = np.zeros(n * ld); yLt = np.zeros(n * ld)
        mM, mL = prog.mv['M'], prog.mv['L']
        lib.call("db_pencil_matvec", n, S, ld, E.ptr(mono), E.ptr(Xt), E.ptr(mM[0]), E.ptr(mM[1]), E.ptr(mM[2]), E.ptr(mM[3]), E.ptr(yMt),
                 E.ptr(mL[0]), E.ptr(mL[1]), E.ptr(mL[2]), E.ptr(mL[3]), E.ptr(yLt), None)
        yM, yL = pi.from_tiles(yMt, n, ld), pi.from_tiles(yLt, n, ld)
        assert np.allclose(yM[:, :S], pi.matvec(prog, 'M', Xref), rtol=1e-13, atol=1e-13)
        assert np.allclose(yL[:, :S], pi.matvec(prog, 'L', Xref), rtol=1e-13, atol=1e-12)
        F = np.zeros((n, ld)); F[:, :S] = rng.standard_normal((n, S))
        Ft = pi.to_tiles(F)
        lc = dlib.LinComb(); lc.nvec = 3
        for j, (v, c) in enumerate(((yMt, 1.0), (Ft, 0.3), (yLt, -0.7))):
            lc.vec[j] = v.ctypes.data; lc.coef[j] = c
        xst = np.zeros(n * ld)
        lib.call("db_pencil_solve", E.ptr(LUt), n, S, ld, E.ptr(_i32(prog.prog)), prog.n_fwd, prog.nE, C.byref(lc), E.ptr(xst), None)
        xs = pi.from_tiles(xst, n, ld)
        rhs = yM + 0.3 * F - 0.7 * yL
        for s in range(0, S, max(1, S // 3)):
            A = batch.matrix((a0, b0), batch.groups[s]).tocsc()
            ref = spsolve(A, rhs[:, s])
            assert np.abs(xs[:, s] - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
        lib.call("db_pencil_scatter", E.ptr(Xt), E.ptr(state_out), S, n, maps.line_base.shape[1], max_len, E.ptr(np.ascontiguousarray(maps.line_base[0])), E.ptr(kind_ext),
                 E.ptr(_i32(maps.line_ptr)), E.ptr(_i32(maps.line_pos)), E.ptr(sys_off), ld, None)
        tmp = np.zeros(var_arena.size); pi.scatter(maps, np.ones((n, S)), tmp); covered |= tmp > 0
    # scatter(gather(state)) reproduces every valid state entry; invalid (sin 0) slots are never touched
    assert np.array_equal(state_out[covered], state[covered])
    nvalid = sum(int(c.valid_cols.sum()) * len(c.groups) for c in builder.classes.values())
    assert covered.sum() == nvalid


def test_pointwise_program():
    lib = E.emu()
    rng = np.random.default_rng(2)
    for npts in (1000, 999):        # vectorised (even) and scalar (odd) kernels
        _check_pointwise(lib, rng, npts)


def _check_pointwise(lib, rng, npts):
    nin = 7
    x = rng.standard_normal((nin, npts))
    # out0 = -x0*x3 - x1*x4 - x2*x5 ; out1 = 2*x6*x6*x0 + 0.5*x1
    term_ptr = _i32([0, 3, 5]); coef = np.array([-1.0, -1.0, -1.0, 2.0, 0.5])
    fac_ptr = _i32([0, 2, 4, 6, 9, 10]); fac = _i32([0, 3, 1, 4, 2, 5, 6, 6, 0, 1])
    out = np.zeros((2, npts))
    lib.call("db_pointwise", E.ptr(x), E.ptr(out), npts, nin, 2, E.ptr(term_ptr), E.ptr(coef), E.ptr(fac_ptr), E.ptr(fac), len(fac), None)
    assert np.allclose(out[0], -x[0] * x[3] - x[1] * x[4] - x[2] * x[5], rtol=1e-15, atol=1e-15)
    assert np.allclose(out[1], 2 * x[6] * x[6] * x[0] + 0.5 * x[1], rtol=1e-15, atol=1e-15)


def test_mmt_and_transposes(golden):
    lib = E.emu()
    rng = np.random.default_rng(4)
    mat = rng.standard_normal((70, 45)); x = rng.standard_normal((3, 45, 67)); out = np.zeros((3, 70, 67))
    lib.call("db_mmt_apply", E.ptr(mat), 70, 45, E.ptr(x), E.ptr(out), 3, 67, None)
    assert np.allclose(out, np.einsum('ij,ojr->oir', mat, x), rtol=1e-13, atol=1e-13)
    # distributed transpose emulated for P ranks: pack -> exchange blocks -> unpack == global transpose of the split
    P, B, n1, n2, n3 = 4, 2, 8, 12, 5
    G = rng.standard_normal((B, n1, n2, n3))
    n1b, n2b = n1 // P, n2 // P
    sends = []
    for r in range(P):
        a = np.ascontiguousarray(G[:, r * n1b:(r + 1) * n1b])
        s = np.zeros(a.size); lib.call("db_transpose_pack", E.ptr(a), E.ptr(s), B, n1b, n2, n3, P, None)
        sends.append(s.reshape(P, -1))
    for r in range(P):
        recv = np.ascontiguousarray(np.stack([sends[src][r] for src in range(P)]))
        out = np.zeros((B, n1, n2b, n3)); lib.call("db_transpose_unpack", E.ptr(recv), E.ptr(out), B, n1, n2b, n3, P, None)
        assert np.array_equal(out, G[:, :, r * n2b:(r + 1) * n2b])
    sends = []
    for r in range(P):
        a = np.ascontiguousarray(G[:, :, r * n2b:(r + 1) * n2b])
        s = np.zeros(a.size); lib.call("db_transpose_pack_rev", E.ptr(a), E.ptr(s), B, n1, n2b, n3, P, None)
        sends.append(s.reshape(P, -1))
    for r in range(P):
        recv = np.ascontiguousarray(np.stack([sends[src][r] for src in range(P)]))
        out = np.zeros((B, n1b, n2, n3)); lib.call("db_transpose_unpack_rev", E.ptr(recv), E.ptr(out), B, n1b, n2, n3, P, None)
        assert np.array_equal(out, G[:, r * n1b:(r + 1) * n1b])
    # absmax
    v = rng.standard_normal(5000); v[1234] = -9.5; res = np.zeros(1)
    lib.call("db_absmax", E.ptr(v), v.size, E.ptr(res), None)
    assert res[0] == 9.5


def test_merged_sign_equivalent_components_share_one_factorisation():
    """pencils.merge_sign_equivalent: the cos / -sin parity blocks of a pencil are D1 A D2 images of one another (D = +-1), so
    they are stored sign-transformed and solved as several right-hand sides of ONE factorisation.  Check (a) the claimed
    relation A_c = D1_c A_0 D2_c entry by entry on numerical pencil matrices, (b) the fused kernels (gather with signs ->
    mat-vec -> solve on 64 * nrhs-thread CTAs -> scatter) against a sparse solve of every member's OWN matrix."""
    from dedalus_b200.solvers import BatchSet
    E.install()
    try:
        pb = examples.rayleigh_benard(dim=3, Nh=8, Nz=12)
        import dedalus_b200 as d3
        solver = pb['problem'].build_solver(d3.RK222)
        a0, b0 = 1.0, 0.01 * GAMMA
        solver._init_device()
        solver._prepare_batches(a0, b0)
        bs = solver.bset
        assert max(db.R for db in bs.items) == 4 and sum(db.R for db in bs.items) == 15      # 15 components, 8 factorisation sets
        bs.factor(0, a0, b0)
        import torch
        rng = np.random.default_rng(3)
        state = torch.from_numpy(rng.standard_normal(solver.var_arena.size))
        feq = torch.from_numpy(rng.standard_normal(solver.eq_arena.size))
        bs.move(0, True, solver.slot_X, state)                 # X   <- D2 state
        bs.move(1, True, solver.slot_F[0], feq)                # F   <- D1 f
        bs.matvec(solver.slot_X, solver.slot_MX0, solver.slot_LX[0])
        bs.solve(0, solver.slot_LX[1], [(solver.slot_F[0], 1.0), (solver.slot_MX0, 0.5)])
        out = torch.zeros_like(state)
        bs.move(0, False, solver.slot_LX[1], out)              # out <- D2 x
        out = out.numpy(); st = state.numpy(); fe = feq.numpy()
        for db in bs.items:
            batch = db.batch
            for c, mem in enumerate(batch.members):
                cols = mem['cols0'][batch.seq]
                for s in range(0, batch.S, max(1, batch.S // 2)):
                    g = batch.groups[s]
                    full = lambda name: batch.builder.class_matrix(batch.cls, name, g, restrict=False).tocsr()
                    Mc, Lc = (full(nm)[mem['rows']][:, cols] for nm in ('M', 'L'))
                    M0, L0 = batch.matrix('M', g), batch.matrix('L', g)
                    D1 = sparse.diags(mem['d1']); D2 = sparse.diags(mem['d2'][batch.seq])
                    assert abs(Mc - D1 @ M0 @ D2).max() < 1e-14 and abs(Lc - D1 @ L0 @ D2).max() < 1e-12
                    # this member's own unknowns / right-hand side picked from the arenas in natural order
                    maps_c = line_maps(batch, solver.var_arena, 'cols'); maps_r = line_maps(batch, solver.eq_arena, 'rows')
                    x_own = pi.gather(maps_c, st, batch.n, batch.S, member=c)[:, s] * mem['d2'][batch.seq]
                    f_own = pi.gather(maps_r, fe, batch.n, batch.S, member=c)[:, s] * mem['d1']
                    rhs = f_own + 0.5 * (Mc @ x_own)
                    ref = spsolve((a0 * Mc + b0 * Lc).tocsc(), rhs)
                    got = pi.gather(maps_c, out, batch.n, batch.S, member=c)[:, s] * mem['d2'][batch.seq]
                    assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max()), (batch.cls.zero_axes, c, s)
    finally:
        E.uninstall()
