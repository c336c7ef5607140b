#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
run() { name=$1; shift
  env "$@" timeout 600 $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1])
    print("$name ms/step", round(d['ms_per_step'],2), {k: (round(v['ms_per_step'],2), round(v['gbps'])) for k, v in d['kernels'].items() if 'axis2' in k})
except Exception as e: print("$name failed", e); print(open('gpurun_out/ab_$name.err').read()[-800:])
PY
}
run base X=1
run fused DB_CHEB_FUSED_SCAN=1
run fusedP8 DB_CHEB_FUSED_SCAN=1 DB_CHEB_P=8
run fusedP2 DB_CHEB_FUSED_SCAN=1 DB_CHEB_P=2
timeout 600 python -m pytest tests/test_gpu_0_transforms.py tests/test_gpu_1_kernels.py -m gpu -q -x -k "swsh or dense_matrix" 2>&1 | tail -3
