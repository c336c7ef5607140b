"""Timed CPU baseline: the oracle's Rayleigh-Benard step (oracle/rb_oracle.py, the reference's algorithm restated
with scipy: pocketfft transforms with the Dedalus packing, per-pencil CSR mat-vecs and SuperLU solves) on a BOUNDED
SAMPLE of the benchmark workload.  TEST / BENCH INFRASTRUCTURE (bench.py cpu_baseline and --impl reference only).

The full 256^3 step costs minutes on a CPU, so each phase is timed on a fraction of its work items and scaled:
  * transforms: every axis pass on a slab holding 1/f of that pass's lines (full line length, all host threads
    through scipy.fft `workers`);
  * grid products: the 19-array pointwise expression on a slab;
  * pencil work (gather, M.x, L.x, RHS combination, SuperLU solve, scatter): `n_pencils` of the (N/2)^2 pencils,
    spread over a process pool (factorisations are built before the clock starts: the reference also reuses them
    while dt is constant, core/timesteppers.py:577-583).
"""
import os, time
import numpy as np
import scipy.fft
from . import rb_oracle
from . import transforms_oracle as T
from . import imex


def _pencil_worker(args):
    dim, Nh, Nz, Ra, groups, dt, reps = args
    # one thread per worker process: the pool already uses every core, and 128 processes x 128 BLAS / OpenMP threads
    # made a single sample take minutes on the GPU box's host
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    orc = rb_oracle.RBOracle(dim, Nh, Nz, Ra)
    tab = imex.RK["RK222"]; A, H = tab['A'], tab['H']
    rng = np.random.default_rng(0)
    data = []
    for grp in groups:
        M, L, vr, vc, voff, eoff = orc.pencil_matrices(grp)
        x = rng.standard_normal(M.shape[1]) * vc
        orc.pencils[grp] = (M, L, vr, vc, voff, eoff)
        orc._solve(grp, (dt, H[1, 1]), 1.0, dt * H[1, 1], np.zeros(M.shape[0]))      # factorise (untimed)
        data.append((grp, x, rng.standard_normal(M.shape[0]) * vr, rng.standard_normal(M.shape[0]) * vr))
    st = orc.new_state(np.zeros(orc.cshape))
    rep_times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        for grp, x, f1, f2 in data:
            M, L, vr, vc, voff, eoff = orc.pencils[grp]
            # stage 1
            xg = orc.gather_state(grp, st, voff)
            mx0 = M @ x; lx0 = L @ x
            rhs = mx0 + dt * A[1, 0] * f1 - dt * H[1, 0] * lx0
            x1 = orc._solve(grp, (dt, H[1, 1]), 1.0, dt * H[1, 1], rhs)
            orc.scatter_state(grp, x1, st, voff)
            # stage 2
            xg = orc.gather_state(grp, st, voff)
            lx1 = L @ x1
            rhs = mx0 + dt * (A[2, 0] * f1 + A[2, 1] * f2) - dt * (H[2, 0] * lx0 + H[2, 1] * lx1)
            x2 = orc._solve(grp, (dt, H[2, 2]), 1.0, dt * H[2, 2], rhs)
            orc.scatter_state(grp, x2, st, voff)
        rep_times.append(time.perf_counter() - t0)
    return rep_times


def sampled_step(dim=3, N=256, Ra=1e6, dt=0.0025, cores=None, n_pencils=None, slab=None, reps=1, pool_timeout=150.0):
    """Estimated wall time of ONE RK222 step of dim-D RB at N^dim on this host, from a bounded sample.  `reps` repeats
    the timed pencil work on the same (untimed) setup -- matrices and factorisations are built once -- and the result
    carries one estimate per repetition in `steps_per_sec_list`; `steps_per_sec` is their median.  If the process pool
    does not finish within `pool_timeout` seconds it is abandoned and a small in-process sample is used instead, so a
    call always returns within a few minutes."""
    cores = cores or os.cpu_count() or 1
    Nh = Nz = N
    G = int(1.5 * N)
    h = dim - 1
    rng = np.random.default_rng(1)
    slab = slab or max(2, min(N, 8))
    # ---- transforms: time one backward and one forward pass per axis on slabs
    def timeit(fn, reps=2):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps
    t_tr_b = t_tr_f = 0.0
    with scipy.fft.set_workers(cores):
        if dim == 3:
            # z pass: lines = Nx*Ny ; sample (slab, Ny, Nz)
            cz = rng.standard_normal((slab, N, N)); frac = slab / N
            tzb = timeit(lambda: T.cheb_backward_fft(cz, G, 2, 0.5, 0.5)) / frac
            gz = rng.standard_normal((slab, N, G))
            tzf = timeit(lambda: T.cheb_forward_fft(gz, N, 2, 1.5, 1.5)) / frac
            # y pass: lines = Nx*Gz ; sample (slab, Ny, Gz)
            cy = rng.standard_normal((slab, N, G))
            tyb = timeit(lambda: T.rf_backward_fft(cy, G, 1)) / frac
            gy = rng.standard_normal((slab, G, G))
            tyf = timeit(lambda: T.rf_forward_fft(gy, N, 1)) / frac
            # x pass: lines = Gy*Gz ; sample (Nx, slab, Gz)
            cx = rng.standard_normal((N, slab, G)); fracx = slab / G
            txb = timeit(lambda: T.rf_backward_fft(cx, G, 0)) / fracx
            gx = rng.standard_normal((G, slab, G))
            txf = timeit(lambda: T.rf_forward_fft(gx, N, 0)) / fracx
            t_tr_b, t_tr_f = tzb + tyb + txb, tzf + tyf + txf
            npts_frac = slab / G
            gin = [rng.standard_normal((slab, G, G)) for _ in range(6)]
        else:
            cz = rng.standard_normal((N, N))
            tzb = timeit(lambda: T.cheb_backward_fft(cz, G, 1, 0.5, 0.5)); gz = rng.standard_normal((N, G))
            tzf = timeit(lambda: T.cheb_forward_fft(gz, N, 1, 1.5, 1.5))
            cx = rng.standard_normal((N, G)); txb = timeit(lambda: T.rf_backward_fft(cx, G, 0))
            gx = rng.standard_normal((G, G)); txf = timeit(lambda: T.rf_forward_fft(gx, N, 0))
            t_tr_b, t_tr_f = tzb + txb, tzf + txf
            npts_frac = 1.0
            gin = [rng.standard_normal((G, G)) for _ in range(6)]
    n_bwd = dim + dim + dim * dim          # u, grad b, grad u  (reference evaluator: 15 scalar transforms in 3-D)
    n_fwd = 1 + dim
    # ---- pointwise: -u.grad(b) and -u.grad(u_j): dim*(1+dim) products
    def pw():
        acc = 0
        for j in range(1 + dim):
            acc = -(gin[0] * gin[3] + gin[1] * gin[4] + gin[2] * gin[5])
        return acc
    t_pw = timeit(pw) / npts_frac
    t_rhs_stage = n_bwd * t_tr_b + n_fwd * t_tr_f + t_pw
    # ---- pencil work on a sample of pencils, process pool
    groups_all = [(i, j) for i in range(N // 2) for j in range(N // 2)] if dim == 3 else [(i,) for i in range(N // 2)]
    n_pencils = n_pencils or max(cores, min(len(groups_all), 2 * cores))
    idx = np.linspace(0, len(groups_all) - 1, n_pencils).astype(int)
    sample = [groups_all[i] for i in idx]
    chunks = [sample[i::cores] for i in range(cores) if sample[i::cores]]
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    used_procs = len(chunks)
    pool = ctx.Pool(len(chunks))
    try:
        times = pool.map_async(_pencil_worker, [(dim, Nh, Nz, Ra, ch, dt, reps) for ch in chunks]).get(timeout=pool_timeout)
        pool.close()
    except mp.TimeoutError:
        pool.terminate()
        n_pencils = min(4, len(groups_all)); used_procs = 1
        idx = np.linspace(0, len(groups_all) - 1, n_pencils).astype(int)
        times = [_pencil_worker((dim, Nh, Nz, Ra, [groups_all[i] for i in idx], dt, reps))]
        cores_eff = 1
        # one process handled the whole sample: scale as if `cores` such processes shared the pencils
        times = [[t / cores for t in times[0]]]
    finally:
        pool.join()
    # per repetition: wall time of the slowest worker for its share
    t_pencil_reps = [max(w[r] for w in times) * len(groups_all) / n_pencils for r in range(reps)]
    t_steps = [2 * t_rhs_stage + tp for tp in t_pencil_reps]
    t_step = float(np.median(t_steps)); t_pencil_step = float(np.median(t_pencil_reps))
    return dict(step_seconds=t_step, steps_per_sec=1.0 / t_step, steps_per_sec_list=[1.0 / t for t in t_steps], cores=cores,
                rhs_seconds=2 * t_rhs_stage, pencil_seconds=t_pencil_step,
                sample=(f"transform passes on {slab}-plane slabs (1/{N // slab} of z/y lines, 1/{G // slab} of x lines), "
                        f"products on a {slab}-plane slab, {n_pencils} of {len(groups_all)} pencils over {used_procs} processes; "
                        f"scaled to the full {N}^{dim} step; SuperLU factorisations excluded (constant dt)"))
