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
    """Mirrors the kernels: chunks of 8 entries, plain x values preloaded at chunk start, FRESH entries served from the
    three most recently completed rows or re-read."""
    END, SKIP, FRESH_REG, FRESH_MEM, MASK, PF = -1, -2, 1 << 30, 1 << 29, (1 << 29) - 1, 8
    n, ld = prog.n, prog.ld
    y = np.array(rhs, dtype=float, copy=True)
    for sec0, sec1, step in ((0, prog.n_fwd, +1), (prog.n_fwd, prog.nE, -1)):
        row = 0 if step > 0 else n - 1
        acc = y[row].copy()
        last = [None, None, None]
        for e0 in range(sec0, sec1, PF):
            codes = [int(prog.prog[e]) for e in range(e0, e0 + PF)]
            pre = [y[c // ld].copy() if (c >= 0 and not (c & (FRESH_REG | FRESH_MEM))) else None for c in codes]
            for j, c in enumerate(codes):
                e = e0 + j
                if c >= 0:
                    if c & FRESH_REG:
                        xv = last[(c & 3) - 1]
                    elif c & FRESH_MEM:
                        xv = y[(c & MASK) // ld]
                    else:
                        xv = pre[j]
                    acc = acc - LU[e] * xv
                elif c == END:
                    val = acc if step > 0 else acc * LU[e]
                    y[row] = val
                    last = [val.copy(), last[0], last[1]]
                    row += step
                    if 0 <= row < n:
                        acc = y[row].copy()
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
