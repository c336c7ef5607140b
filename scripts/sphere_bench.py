"""Timing of BASELINE config 4 (spherical shallow water, Nphi x Ntheta = 512 x 256, Lmax = 254, fp64, RK222) on one GPU.
Not the round's bench line (bench.py measures config 3); prints one JSON line with steps/s, launches and the per-kernel
table (CUDA events around every C-ABI launch, same accounting as bench.py).  2-D sphere steps are latency-bound (state
1.5 MB): steps/s and launches per step are the meaningful figures, SURVEY section 8d."""
import argparse, json, sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="shallow_water", choices=["shallow_water", "shell_convection"])
    ap.add_argument("--nphi", type=int, default=512)
    ap.add_argument("--ntheta", type=int, default=256)
    ap.add_argument("--nr", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    import torch
    import dedalus_b200 as d3
    from dedalus_b200 import examples
    from dedalus_b200.lib import get_lib
    torch.cuda.set_device(0)
    t0 = time.time()
    if args.problem == "shell_convection":          # config 5's problem (single GPU): python scripts/sphere_bench.py --problem shell_convection --nphi 256 --ntheta 128
        sw = examples.shell_convection(args.nphi, args.ntheta, args.nr)
        solver = sw['problem'].build_solver(d3.SBDF2)
        examples.shell_convection_initial_condition(sw['b'], sw['shell'], sw['Ri'], sw['Ro'])
        sw.update(h=sw['b'], basis=sw['shell'].sphere_basis, timestep=0.05)
    else:
        sw = examples.shallow_water(args.nphi, args.ntheta)
        solver = sw['problem'].build_solver(d3.RK222)
        examples.shallow_water_initial_condition(sw['u'], sw['h'], sw['basis'], sw['units'])
    dt = sw['timestep']
    for _ in range(max(args.warmup, 3)):
        solver.step(dt)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    lib = get_lib()
    # un-instrumented timing first (the per-launch events add host work to a launch-bound step)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.launches
    e0.record()
    for _ in range(args.steps):
        solver.step(dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = lib.launches - l0
    solver.prof = []
    for _ in range(args.steps):
        solver.step(dt)
    torch.cuda.synchronize()
    agg = {}
    for name, a, b, nbytes in solver.prof:
        d = agg.setdefault(name, dict(ms=0.0, bytes=0, launches=0))
        d['ms'] += a.elapsed_time(b); d['bytes'] += nbytes; d['launches'] += 1
    solver.prof = None
    kernels = {k: dict(ms_per_step=d['ms'] / args.steps, launches_per_step=d['launches'] / args.steps,
                       gbps=d['bytes'] / (d['ms'] * 1e-3) / 1e9 if d['ms'] > 0 else None)
               for k, d in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])}
    u, h = sw['u']['c'], sw['h']['c']
    print(json.dumps(dict(metric=f"timesteps/sec {args.problem} {args.nphi}x{args.ntheta}" + (f"x{args.nr}" if args.problem == "shell_convection" else "") +
                                 f" (Lmax={sw['basis'].Lmax}) fp64, 1 B200",
                          value=args.steps / (ms * 1e-3), unit="steps/s", ms_per_step=ms / args.steps, steps=args.steps,
                          gpu_launches_per_step=launches / args.steps, setup_seconds=setup_s, kernels=kernels,
                          pencil_systems=solver.bset.nsys, band=(getattr(solver.bset, 'kl', None), getattr(solver.bset, 'ku', None)),
                          max_n=getattr(solver.bset, 'max_n', getattr(solver.bset, 'n', None)),
                          factor_backward_error=solver.bset.last_verify, finite=bool(np_finite(u) and np_finite(h)),
                          steps_taken=int(solver.iteration))))


def np_finite(a):
    import numpy as np
    return bool(np.isfinite(a).all())


if __name__ == "__main__":
    main()
