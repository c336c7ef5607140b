"""Time the fused pencil solve on ONE GPU for a slab-shaped problem (Nx x N x N): the per-GPU share of the N^3 benchmark
on P = N / Nx ranks, i.e. the strong-scaling regime of k_batches_solve_* without paying for P GPUs.
usage: python scripts/solve_microbench.py [Nx] [N]   (kernel variant via the DB_SOLVE_* environment switches)"""
import sys, os, json, time
sys.path.insert(0, ".")
import numpy as np, torch
import dedalus_b200 as d3
from dedalus_b200 import examples

Nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
pb = examples.rayleigh_benard(dim=3, Nh=N, Nz=N, Nx=Nx)
solver = pb['problem'].build_solver(d3.RK222)
examples.rayleigh_benard_initial_condition(pb['b'], pb['bases'], pb['Lz'])
dt = 1e-2 * 64.0 / N
solver.step(dt)                         # builds, factorises, verifies
bs = solver.bset
tiles = bs.blocks['solve']
terms = [(solver.slot_MX0, 1.0), (solver.slot_F[0], 0.3), (solver.slot_LX[0], -0.1), (solver.slot_F[1], 0.2)]
for _ in range(3):
    bs.solve(0, solver.slot_X, terms)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    bs.solve(0, solver.slot_X, terms)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# correctness of whatever variant ran: residual probe
res = bs.probe(0, 1.0, dt * float(solver.timestepper_class.H[1, 1]), (solver.slot_F[0], solver.slot_X, solver.slot_MX0, solver.slot_LX[0]), seed=5)
worst = max(float(r.max()) for r in res if r.size)
nbytes = 8 * (bs.sum_ES + bs.sum_nS * (len(terms) + 1))
print(json.dumps(dict(Nx=Nx, N=N, tiles=tiles, ms_per_solve=ms, algorithmic_gbps=nbytes / ms / 1e6, backward_error=worst,
                      env={k: v for k, v in os.environ.items() if k.startswith("DB_SOLVE")})))
