#!/bin/bash
# A/B of the triangular-solve kernels at 256^3 on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --size 256 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1])
k = d['kernels']
print("$name", "steps/s", round(d['value'], 2), "solve ms", round(k['pencil_solve']['ms_per_step'], 2), "GB/s", round(k['pencil_solve']['gbps']), "matvec", round(k['pencil_matvec']['ms_per_step'], 2), "pointwise", round(k['pointwise']['ms_per_step'], 2))
PY
}
run ring3 DB_SOLVE_IMPL=ring
run flat2 DB_SOLVE_STAGES=2
run flat3 DB_SOLVE_STAGES=3
run flat4 DB_SOLVE_STAGES=4
