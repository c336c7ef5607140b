"""Shared by the emulated (small) and GPU (256^3) tests: size-independent property of the pencil solve -- for a random
right-hand side b, x = solve(b) through the fused kernels satisfies (M + b0 L) x = b when multiplied back through the fused
mat-vec kernel, for every system of every batch (the analogue of an encode -> decode round trip for S3 / S4)."""
import numpy as np


def solve_residual(solver, dt, seed=0):
    import torch
    cls = solver.timestepper_class
    assert cls.kind == "rk"
    b0 = dt * float(cls.H[1, 1])
    bs = solver.bset
    s_b, s_x, s_m, s_l = solver.slot_F[0], solver.slot_LX[1], solver.slot_MX0, solver.slot_LX[0]
    gen = torch.Generator(device='cpu'); gen.manual_seed(seed)
    for db in bs.items:
        v = db.vecs[s_b]
        v.copy_(torch.randn(v.numel(), generator=gen, dtype=torch.float64).to(v.device))
    bs.solve(solver._stage_lu[0], s_x, [(s_b, 1.0)])
    bs.matvec(s_x, s_m, s_l)
    worst = 0.0
    for db in bs.items:
        n, S, ld = db.n, db.S, db.ld
        view = lambda t: t.view(ld // 64, n, 64)
        b, x, mx, lx = (view(db.vecs[s]) for s in (s_b, s_x, s_m, s_l))
        valid = (torch.arange(ld, device=b.device).view(ld // 64, 1, 64) < S)
        r = (mx + b0 * lx - b) * valid
        scale = float(((b.abs() + mx.abs() + b0 * lx.abs()) * valid).max())
        assert bool(torch.isfinite(x * valid).all())
        worst = max(worst, float(r.abs().max()) / scale)
    return worst
