#!/bin/bash
# generic A/B harness: each argument "name:ENV=VAL,ENV2=VAL2" is one 256^3 bench run
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}; envs=${envs//,/ }
  env $envs timeout 600 python bench.py --size 256 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1])
k = d['kernels']
print("$name", "steps/s", round(d['value'], 2), " ".join(f"{n.replace('transform_','t_').replace('pencil_','')}={v['ms_per_step']:.2f}" for n, v in k.items()))
PY
done
