"""TEST INFRASTRUCTURE: numpy interpreters of the batch programs built by dedalus_b200/pencils.py.
They execute exactly what csrc/pencil.cu executes per thread (one system per thread, SoA arrays
[entry][system]) and are used by the CPU tests to validate the programs against dense/sparse solves."""
import numpy as np


def assemble(prog, asm):
    ptr, mono, val = asm
    LU = np.zeros((prog.nE, prog.S))
    for e in range(prog.nE):
        for t in range(ptr[e], ptr[e + 1]):
            LU[e] += val[t] * prog.mono_vals[mono[t]]
    return LU


def factor(prog, LU):
    n = prog.n
    dp = 0
    for k in range(n):
        d = prog.diag_eid[k]
        inv = 1.0 / LU[d]
        LU[d] = inv
        us = prog.fu_eid[prog.fu_ptr[k]:prog.fu_ptr[k + 1]]
        for a in range(prog.fl_ptr[k], prog.fl_ptr[k + 1]):
            le = prog.fl_eid[a]
            l = LU[le] * inv
            LU[le] = l
            for ue in us:
                LU[prog.fd_eid[dp]] -= l * LU[ue]
                dp += 1
    return LU


def solve(prog, LU, rhs, pipelined=False):
    """Mirrors csrc/pencil.cu k_batches_solve_flat (pipelined=True: k_batches_solve_pipe, whose gathers are issued one chunk
    early and which re-reads the entries of control word 35 instead of 34): x <- rhs; per 16-entry chunk the control block's 16 offsets are
    gathered in one burst, entries flagged in maskF are re-read right before use, row-boundary entries store the
    accumulator into the row being left (backward: times the reciprocal pivot) and continue from the gathered start
    value of the row being entered."""
    CH = 16
    ld = prog.tile
    y = np.array(rhs, dtype=float, copy=True)
    ctrl = prog.ctrl
    nfwd = prog.n_fwd // CH
    acc = None
    nchunks = prog.nE // CH
    gather = lambda q: [y[g // ld].copy() for g in ctrl[q, :16]]
    nxt = gather(0) if pipelined else None
    for q in range(nchunks):
        forward = q < nfwd
        goff, foff = ctrl[q, :16], ctrl[q, 16:32]
        maskE, maskB, maskF = (int(ctrl[q, 32]) & 0xFFFF, int(ctrl[q, 33]) & 0xFFFF, int(ctrl[q, 35 if pipelined else 34]) & 0xFFFF)
        if pipelined:
            xv = nxt
            nxt = gather(q + 1) if q + 1 < nchunks else None          # issued before chunk q is consumed
        else:
            xv = gather(q)
        for j in range(CH):
            e = q * CH + j
            if maskF >> j & 1:
                xv[j] = y[goff[j] // ld].copy()
            if acc is None:
                acc = np.zeros_like(xv[j])
            acc_a = acc - LU[e] * xv[j]
            val = acc if forward else acc * LU[e]
            if maskE >> j & 1:
                y[foff[j] // ld] = val
            acc = xv[j] if (maskB >> j & 1) else acc_a
    return y


def solve_deep(prog, LU, rhs, D=None, RRN=None):
    """Mirrors csrc/pencil.cu k_batches_solve_deep: the 16 gathers of chunk q are issued D chunks ahead (they see x as it
    was before chunk q - D was consumed); entries in maskR take their value from the ring of the last RRN finished rows
    (slot from the packed rslot words), entries in maskG re-read global memory right before use."""
    from dedalus_b200.pencils import SOLVE_DEEP_D, SOLVE_DEEP_RING
    D = SOLVE_DEEP_D if D is None else D
    RRN = SOLVE_DEEP_RING if RRN is None else RRN
    CH, ld = 16, prog.tile
    y = np.array(rhs, dtype=float, copy=True)
    ctrl = prog.ctrl
    nchunks, nfwd = prog.nE // CH, prog.n_fwd // CH
    gather = lambda q: [y[g // ld].copy() for g in ctrl[q, :16]]
    inflight = {q: gather(q) for q in range(min(D, nchunks))}          # issued before chunk 0 is consumed
    ring = [None] * RRN
    cnt = 0
    acc = np.zeros_like(y[0])
    for q in range(nchunks):
        if q + D < nchunks:
            inflight[q + D] = gather(q + D)                              # top of iteration q
        xv = inflight.pop(q)
        forward = q < nfwd
        goff, foff = ctrl[q, :16], ctrl[q, 16:32]
        maskE, maskB, maskR, maskG = (int(ctrl[q, w]) & 0xFFFF for w in (32, 33, 36, 37))
        for j in range(CH):
            e = q * CH + j
            if maskR >> j & 1:
                slot = (int(ctrl[q, 38 + j // 6]) >> (5 * (j % 6))) & 31
                xv[j] = ring[slot].copy()
            if maskG >> j & 1:
                xv[j] = y[goff[j] // ld].copy()
            acc_a = acc - LU[e] * xv[j]
            val = acc if forward else acc * LU[e]
            if maskE >> j & 1:
                y[foff[j] // ld] = val
                ring[cnt % RRN] = val.copy()
                cnt += 1
            acc = xv[j] if (maskB >> j & 1) else acc_a
    return y


def matvec(prog, name, x):
    ptr, col, mono, val = prog.mv[name]
    y = np.zeros_like(x)
    for i in range(prog.n):
        for t in range(ptr[i], ptr[i + 1]):
            y[i] += val[t] * prog.mono_vals[mono[t]] * x[col[t]]
    return y


def gather(maps, arena, n, S, member=0):
    """Member `member` of a (possibly merged) batch: line_base / line_sign are (R, nlines)."""
    X = np.zeros((n, S))
    base, sign = np.atleast_2d(maps.line_base)[member], np.atleast_2d(maps.line_sign)[member]
    for q in range(len(base)):
        for m in range(maps.line_len[q]):
            pos = maps.line_pos[maps.line_ptr[q] + m]
            X[pos] = sign[q] * arena[base[q] + maps.sys_off[maps.line_kind[q]] + m]
    return X


def scatter(maps, X, arena, member=0):
    base, sign = np.atleast_2d(maps.line_base)[member], np.atleast_2d(maps.line_sign)[member]
    for q in range(len(base)):
        for m in range(maps.line_len[q]):
            pos = maps.line_pos[maps.line_ptr[q] + m]
            arena[base[q] + maps.sys_off[maps.line_kind[q]] + m] = sign[q] * X[pos]


def to_tiles(a, tile=64):
    """(rows, ld) row-major array -> flat tile-major storage used by the kernels (ld multiple of `tile`)."""
    rows, ld = a.shape
    return np.ascontiguousarray(a.reshape(rows, ld // tile, tile).transpose(1, 0, 2)).reshape(-1)


def from_tiles(flat, rows, ld, tile=64):
    return np.ascontiguousarray(flat.reshape(ld // tile, rows, tile).transpose(1, 0, 2)).reshape(rows, ld)
