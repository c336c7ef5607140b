"""Shared by the emulated (small) and GPU (256^3) tests: size-independent property of the pencil solve -- for a random
right-hand side b, x = solve(b) through the fused kernels satisfies (M + b0 L) x = b when multiplied back through the fused
mat-vec kernel, for every system of every batch (the analogue of an encode -> decode round trip for S3 / S4).  The residual
itself is formed by db_batches_residual (the product's own verification, run after every factorisation); here it is called
with another seed and additionally cross-checked against a torch evaluation of the same expression."""
import numpy as np


def solve_residual(solver, dt, seed=0):
    import torch
    cls = solver.timestepper_class
    assert cls.kind == "rk"
    b0 = dt * float(cls.H[1, 1])
    bs = solver.bset
    s_b, s_x, s_m, s_l = solver.slot_F[0], solver.slot_LX[1], solver.slot_MX0, solver.slot_LX[0]
    res = bs.probe(solver._stage_lu[0], 1.0, b0, (s_b, s_x, s_m, s_l), seed=seed)
    worst = max(float(r.max()) for r in res if r.size)
    # independent evaluation of the same quantity from the vectors the probe left behind
    check = 0.0
    for db in bs.items:
        n, S, ld, R = db.n, db.S, db.ld, db.R
        view = lambda t: t[:R * ld * n].view(R * ld // 64, n, 64)      # (allocated for 4 members; the first R are real)
        b, x, mx, lx = (view(db.vecs[s]) for s in (s_b, s_x, s_m, s_l))
        col = torch.arange(R * ld, device=b.device).view(R * ld // 64, 1, 64) % ld
        valid = col < S
        r = ((mx + b0 * lx - b) * valid).abs().amax(dim=1)
        den = (b.abs() * valid).amax(dim=1) + (mx.abs() * valid).amax(dim=1) + (b0 * lx.abs() * valid).amax(dim=1)
        assert bool(torch.isfinite(x * valid).all())
        check = max(check, float((r / den.clamp_min(1e-300)).max()))
    assert abs(check - worst) <= 1e-3 * worst + 1e-18, (check, worst)
    return worst
