"""Invariants of the host-built programs the fused kernels execute (dedalus_b200/pencils.py solve stream + control blocks,
dedalus_b200/solvers.py mat-vec windows): pure host logic, no kernels."""
import numpy as np, pytest
from dedalus_b200 import examples
from dedalus_b200.pencils import PencilSystemBuilder, build_batches, compile_batch, SOLVE_CTRL_WORDS
from dedalus_b200.solvers import matvec_windows

SKIP = -2**31


@pytest.fixture(scope="module")
def programs():
    pb = examples.rayleigh_benard(dim=3, Nh=8, Nz=24, Rayleigh=1e6)
    builder = PencilSystemBuilder(pb['problem'])
    out = []
    for batch in build_batches(builder):
        for dense in (64, 6):
            out.append(compile_batch(batch, 1.0, 0.01, dense=dense))
    return out


def test_solve_stream_structure(programs):
    for prog in programs:
        code = prog.prog.astype(np.int64)
        assert len(code) == prog.nE and prog.nE % 16 == 0 and prog.n_fwd % 16 == 0
        for a, e, forward in ((0, prog.n_fwd, True), (prog.n_fwd, prog.nE, False)):
            sec = code[a:e]
            assert sec[0] < 0 and sec[0] != SKIP                     # a section starts by entering a row
            cur, final_store = None, {}
            entered = set()
            for pos, c in enumerate(sec):
                if c == SKIP:
                    continue
                if c < 0:
                    if cur is not None:
                        final_store[cur] = pos
                    cur = (-1 - c) // prog.tile
                    entered.add(cur)
                else:
                    col = c // prog.tile
                    assert c % prog.tile == 0 and 0 <= col < prog.n
                    # a column is only read after its row was left for the last time (its value is final)
                    assert col != cur
            assert entered == set(range(prog.n))                      # every row visited
            # dependencies: forward rows read lower-numbered pivots only, backward rows higher-numbered ones
            cur = None
            for c in sec:
                if c == SKIP:
                    continue
                if c < 0:
                    cur = (-1 - c) // prog.tile
                else:
                    assert (c // prog.tile < cur) if forward else (c // prog.tile > cur)


def test_control_blocks_mirror_the_stream(programs):
    for prog in programs:
        ctrl = prog.ctrl
        assert ctrl.shape == (prog.nE // 16, SOLVE_CTRL_WORDS) and ctrl.dtype == np.int32
        code = prog.prog.astype(np.int64)
        for q in range(ctrl.shape[0]):
            mE, mB, mF = (int(ctrl[q, 32 + i]) & 0xFFFF for i in range(3))
            assert mE & ~mB == 0                                       # leaving a row always enters the next one
            for j in range(16):
                c = code[16 * q + j]
                if c == SKIP:
                    assert ctrl[q, j] == 0 and not (mB >> j) & 1
                elif c < 0:
                    assert (mB >> j) & 1 and ctrl[q, j] == -1 - c
                else:
                    assert not (mB >> j) & 1 and ctrl[q, j] == c
        # the late-read mask marks exactly the gathers whose source row is stored earlier in the same chunk
        stored = {}
        for e, c in enumerate(code):
            q, j = divmod(e, 16)
            if e == prog.n_fwd:
                cur = None
            if e == 0:
                cur = None
            if c == SKIP:
                continue
            src = (-1 - c) if c < 0 else c
            late = stored.get(src, -1) >= 16 * q
            assert bool((int(ctrl[q, 34]) >> j) & 1) == late
            if c < 0:
                if cur is not None:
                    stored[cur] = e
                cur = -1 - c


def test_matvec_windows_cover_the_band(programs):
    for prog in programs[::2]:
        win = matvec_windows(prog, 32, 80)
        assert win.shape == (-(-prog.n // 32), 2)
        inside = total = 0
        for rb, (w0, wl) in enumerate(win):
            assert 0 <= w0 and wl <= 80 and w0 + wl <= prog.n
            r0, r1 = 32 * rb, min(prog.n, 32 * rb + 32)
            for k in ('M', 'L'):
                ptr, col = prog.mv[k][0], np.asarray(prog.mv[k][1])
                cols = col[ptr[r0]:ptr[r1]]
                total += cols.size; inside += int(((cols >= w0) & (cols < w0 + wl)).sum())
        assert inside >= 0.6 * total                                 # the window holds the bulk of every block's terms
