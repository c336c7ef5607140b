"""csrc/pencil.cu + pointwise.cu logic under the test-only CPU emulation vs the numpy program interpreters /
direct numpy evaluation."""
import numpy as np, pytest, ctypes as C
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
    batches = build_batches(builder)
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
        lib.call("db_pencil_gather", E.ptr(state), E.ptr(Xt), S, n, len(maps.line_base), max_len, E.ptr(maps.line_base), E.ptr(kind_ext),
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
        lib.call("db_pencil_scatter", E.ptr(Xt), E.ptr(state_out), S, n, len(maps.line_base), max_len, E.ptr(maps.line_base), E.ptr(kind_ext),
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
