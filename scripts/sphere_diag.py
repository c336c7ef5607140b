"""Diagnosis of the banded pencil kernels at BASELINE config 4's size (255 systems, n <= 1530, kl = ku = 7): every kernel
variant (DB_BANDED_MODE 0..3) is run repeatedly; the factors are compared with LAPACK's dgbtrf on the host (values and pivot
rows), the solves with dgbtrs, and both with themselves across repetitions (bitwise reproducibility).  Then config 4 is timed
(scripts/sphere_bench.py accounting).  Prints one JSON line."""
import json, sys, pathlib, time
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    from scipy.linalg import lapack
    import dedalus_b200 as d3
    from dedalus_b200 import examples, sphere
    from dedalus_b200.lib import get_lib
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    torch.cuda.set_device(0)
    sw = examples.shallow_water(512, 256)
    solver = sw['problem'].build_solver(d3.RK222)
    dt = sw['timestep']
    solver._init_device()
    a0, b0 = 1.0, dt * float(d3.RK222.H[1, 1])
    solver._prepare_batches(a0, b0)
    bs = solver.bset
    lib = get_lib()
    kl, ku = bs.kl, bs.ku
    ldf = 2 * kl + ku + 1
    # host reference factorisation of every system
    M, L = bs.M_ab.cpu().numpy(), bs.L_ab.cpu().numpy()
    ld0 = kl + ku + 1
    gen = torch.Generator(device=solver.device); gen.manual_seed(99)
    bvec = torch.zeros_like(bs.vecs[0]).normal_(generator=gen)
    bh = bvec.cpu().numpy()
    ref_lu, ref_piv, ref_x = [], [], []
    for s in bs.systems:
        n, o = s['n'], s['op_off']
        ab0 = (a0 * M[o:o + n * ld0] + b0 * L[o:o + n * ld0]).reshape(n, ld0).T          # (ld0, n) operator storage
        abf = np.zeros((ldf, n), order='F'); abf[kl:, :] = ab0
        lu, piv, info = lapack.dgbtrf(abf, kl, ku)
        x, info2 = lapack.dgbtrs(lu, kl, ku, bh[s['vec_off']:s['vec_off'] + n].copy(), piv)
        ref_lu.append(np.asarray(lu)); ref_piv.append(np.asarray(piv)); ref_x.append(x)
    out = dict(reps=reps, nsys=bs.nsys, kl=kl, max_n=bs.max_n, modes={})
    vec_off = [s['vec_off'] for s in bs.systems]
    for mode in (0, 1, 2, 3):
        lib._raw_db_banded_set_mode(mode)
        rec = dict(verify=[], factor_bitwise_repro=True, solve_bitwise_repro=True, lu_max_rel_err=0.0, piv_mismatch=0, x_max_rel_err=0.0)
        lu0 = x0 = None
        for r in range(reps):
            bs.factor(0, a0, b0)
            lu = bs.lu[0].cpu().numpy().copy(); piv = bs.ipiv[0].cpu().numpy().copy()
            if lu0 is None:
                lu0 = lu
                lo = 0
                for i, s in enumerate(bs.systems):
                    n = s['n']
                    mine = lu[lo:lo + n * ldf].reshape(n, ldf).T
                    rec['lu_max_rel_err'] = max(rec['lu_max_rel_err'], float(np.abs(mine - ref_lu[i]).max() / np.abs(ref_lu[i]).max()))
                    rec['piv_mismatch'] += int((piv[vec_off[i]:vec_off[i] + n] != ref_piv[i]).sum())
                    lo += n * ldf
            elif not np.array_equal(lu, lu0):
                rec['factor_bitwise_repro'] = False
            bs.vecs[1].copy_(bvec)
            bs.solve(0, 2, [(1, 1.0)])
            x = bs.vecs[2].cpu().numpy().copy()
            if x0 is None:
                x0 = x
                for i, s in enumerate(bs.systems):
                    xi = x[vec_off[i]:vec_off[i] + s['n']]
                    rec['x_max_rel_err'] = max(rec['x_max_rel_err'], float(np.abs(xi - ref_x[i]).max() / (np.abs(ref_x[i]).max() + 1e-300)))
            elif not np.array_equal(x, x0):
                rec['solve_bitwise_repro'] = False
            bs.matvec(2, 3, 4)
            res = (a0 * bs.vecs[3] + b0 * bs.vecs[4] - bvec).abs().max() / (bvec.abs().max() + (a0 * bs.vecs[3]).abs().max() + (b0 * bs.vecs[4]).abs().max())
            rec['verify'].append(float(res))
        out['modes'][str(mode)] = rec
    # pick the first variant whose factor and solve agree with LAPACK, time config 4 with it
    good = [m for m, r in out['modes'].items() if max(r['verify']) < 1e-10]
    use = int(good[0]) if good else 0
    lib._raw_db_banded_set_mode(use)
    out['timed_mode'] = use
    sphere.SphereSystems.VERIFY_TOL = 1.0
    examples.shallow_water_initial_condition(sw['u'], sw['h'], sw['basis'], sw['units'])
    for _ in range(3):
        solver.step(dt)
    torch.cuda.synchronize()
    steps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.launches
    e0.record()
    for _ in range(steps):
        solver.step(dt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.launches - l0
    solver.prof = []
    for _ in range(steps):
        solver.step(dt)
    torch.cuda.synchronize()
    agg = {}
    for name, a, b, nbytes in solver.prof:
        d = agg.setdefault(name, dict(ms=0.0, bytes=0, launches=0))
        d['ms'] += a.elapsed_time(b); d['bytes'] += nbytes; d['launches'] += 1
    solver.prof = None
    out['bench'] = dict(steps_per_s=steps / (ms * 1e-3), ms_per_step=ms / steps, launches_per_step=launches / steps,
                        first_step_verify=bs.last_verify,
                        kernels={k: dict(ms_per_step=d['ms'] / steps, launches_per_step=d['launches'] / steps,
                                         gbps=d['bytes'] / (d['ms'] * 1e-3) / 1e9 if d['ms'] > 0 else None)
                                 for k, d in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])},
                        finite=bool(np.isfinite(sw['u']['c']).all() and np.isfinite(sw['h']['c']).all()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
