#!/bin/bash
# round 2, first GPU contact: full GPU suite (no -x) + default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log; tail -15 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench256.json 2> gpurun_out/r2_bench256.err; echo "rc=$?" >> gpurun_out/r2_bench256.err
tail -3 gpurun_out/r2_bench256.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench256.json').read().strip().splitlines()[-1])
print("value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e']['value'], "launches", d['gpu_launches'])
for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s  share {v['share']:.3f}")
PY
