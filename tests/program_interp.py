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


def solve(prog, LU, rhs):
    """Mirrors csrc/pencil.cu k_batches_solve: chunks of 8 entries; plain x values (and the start value of the next row)
    of chunk q are preloaded BEFORE chunk q-1 is computed; FRESH entries come from the three most recently completed rows
    or are re-read at compute time."""
    SKIP, FRESH_REG, FRESH_MEM, MASK, CH = -2**31, 1 << 30, 1 << 29, (1 << 29) - 1, 16
    n, ld = prog.n, prog.tile
    y = np.array(rhs, dtype=float, copy=True)
    for sec0, sec1, forward in ((0, prog.n_fwd, True), (prog.n_fwd, prog.nE, False)):
        nch = (sec1 - sec0) // CH

        def preload(q):
            out = []
            for e in range(sec0 + q * CH, sec0 + (q + 1) * CH):
                c = int(prog.prog[e])
                if c == SKIP:
                    out.append(None)
                elif c < 0:
                    out.append(y[(-1 - c) // ld].copy())          # rhs (forward) / forward result (backward) of the next row
                elif c & (FRESH_REG | FRESH_MEM):
                    out.append(None)
                else:
                    out.append(y[c // ld].copy())
            return out
        cur, acc, last = None, None, [None, None, None]
        DB = 0                                         # x values of chunk q are gathered right before chunk q is computed
        for q in range(nch):
            pre = preload(q)
            for j in range(CH):
                e = sec0 + q * CH + j
                c = int(prog.prog[e])
                if c == SKIP:
                    continue
                if c < 0:
                    if cur is not None:
                        val = acc if forward else acc * LU[e]
                        y[cur] = val
                        last = [val.copy(), last[0], last[1]]
                    cur = (-1 - c) // ld
                    acc = pre[j]
                else:
                    if c & FRESH_REG:
                        xv = last[(c & 3) - 1]
                    elif c & FRESH_MEM:
                        xv = y[(c & MASK) // ld]
                    else:
                        xv = pre[j]
                    acc = acc - LU[e] * xv
    return y


def matvec(prog, name, x):
    ptr, col, mono, val = prog.mv[name]
    y = np.zeros_like(x)
    for i in range(prog.n):
        for t in range(ptr[i], ptr[i + 1]):
            y[i] += val[t] * prog.mono_vals[mono[t]] * x[col[t]]
    return y


def gather(maps, arena, n, S):
    X = np.zeros((n, S))
    for q in range(len(maps.line_base)):
        for m in range(maps.line_len[q]):
            pos = maps.line_pos[maps.line_ptr[q] + m]
            X[pos] = arena[maps.line_base[q] + maps.sys_off[maps.line_kind[q]] + m]
    return X


def scatter(maps, X, arena):
    for q in range(len(maps.line_base)):
        for m in range(maps.line_len[q]):
            pos = maps.line_pos[maps.line_ptr[q] + m]
            arena[maps.line_base[q] + maps.sys_off[maps.line_kind[q]] + m] = X[pos]


def to_tiles(a, tile=64):
    """(rows, ld) row-major array -> flat tile-major storage used by the kernels (ld multiple of `tile`)."""
    rows, ld = a.shape
    return np.ascontiguousarray(a.reshape(rows, ld // tile, tile).transpose(1, 0, 2)).reshape(-1)


def from_tiles(flat, rows, ld, tile=64):
    return np.ascontiguousarray(flat.reshape(ld // tile, rows, tile).transpose(1, 0, 2)).reshape(rows, ld)
