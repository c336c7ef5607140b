"""Static-ordering LU programs vs sparse direct solves of the same pencil systems (reference solver:
matsolvers.py:179-183 SuperLU on each pencil; here scipy spsolve on the assembled matrix)."""
import numpy as np, pytest
from scipy import sparse
from scipy.sparse.linalg import spsolve
from dedalus_b200 import examples
from dedalus_b200.pencils import PencilSystemBuilder, build_batches, compile_batch, assembly_program
import program_interp as pi

GAMMA = (2 - np.sqrt(2)) / 2


@pytest.mark.parametrize("dense", [64, 5])      # 5: the segmented dense-row runs are exercised at test sizes too
@pytest.mark.parametrize("dim,Nh,Nz,dt", [(3, 8, 16, 1e-2), (2, 16, 24, 1e-3), (3, 6, 12, 1e2)])
def test_rb_programs_solve(dim, Nh, Nz, dt, dense):
    pb = examples.rayleigh_benard(dim=dim, Nh=Nh, Nz=Nz)
    builder = PencilSystemBuilder(pb['problem'])
    batches = build_batches(builder, merge=False)      # single-component batches: the per-batch kernels and interpreters
    a0, b0 = 1.0, dt * GAMMA
    rng = np.random.default_rng(0)
    total = 0
    for batch in batches:
        prog = compile_batch(batch, a0, b0, dense=dense)
        asm = assembly_program(batch, prog, a0, b0)
        LU = pi.factor(prog, pi.assemble(prog, asm))
        rhs = rng.standard_normal((prog.n, prog.S))
        x = pi.solve(prog, LU, rhs)
        assert np.allclose(pi.solve(prog, LU, rhs, pipelined=True), x, rtol=0, atol=0)      # one-chunk-early gathers + word-35 re-reads
        assert np.allclose(pi.solve_deep(prog, LU, rhs), x, rtol=0, atol=0)                 # deep prefetch + ring of recent rows
        xm = rng.standard_normal((prog.n, prog.S))
        Mx = pi.matvec(prog, 'M', xm); Lx = pi.matvec(prog, 'L', xm)
        for s in range(0, prog.S, max(1, prog.S // 5)):
            A = batch.matrix((a0, b0), batch.groups[s]).tocsc()
            ref = spsolve(A, rhs[:, s])
            err = np.abs(x[:, s] - ref).max() / np.abs(ref).max()
            assert err < 1e-9, (batch.cls.zero_axes, s, err)
            assert np.allclose(batch.matrix('M', batch.groups[s]) @ xm[:, s], Mx[:, s], rtol=1e-12, atol=1e-12)
            assert np.allclose(batch.matrix('L', batch.groups[s]) @ xm[:, s], Lx[:, s], rtol=1e-12, atol=1e-10)
        total += prog.n * prog.S
    # all valid degrees of freedom are covered exactly once
    ndof = sum(int(c.valid_rows.sum()) * len(c.groups) for c in builder.classes.values())
    assert total == ndof
