#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1])
    print("$name ms/step", round(d['ms_per_step'],2), {k: (round(v['ms_per_step'],2), round(v['gbps'])) for k, v in d['kernels'].items() if 'solve' in k or 'axis2' in k})
except Exception as e: print("$name failed", e); print(open('gpurun_out/ab_$name.err').read()[-1500:])
PY
}
run rt1 DB_SOLVE_RT=1
run rt2 DB_SOLVE_RT=2
run rt4 DB_SOLVE_RT=4
run rt2st4 DB_SOLVE_RT=2 DB_SOLVE_STAGES=4
run rt4st4 DB_SOLVE_RT=4 DB_SOLVE_STAGES=4
run rt4st3 DB_SOLVE_RT=4 DB_SOLVE_STAGES=3
timeout 900 python -m pytest tests/test_gpu_0_transforms.py tests/test_gpu_solver.py -m gpu -q -x -k "cheb or Cheb or tall or known_answer or register" > gpurun_out/r2_pytest_sub.log 2>&1; tail -3 gpurun_out/r2_pytest_sub.log
