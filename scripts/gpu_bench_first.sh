#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 1500 python bench.py --size 256 --steps 10 --warmup 3 > gpurun_out/bench256.json 2> gpurun_out/bench256.err; echo "rc=$?" >> gpurun_out/bench256.err
tail -3 gpurun_out/pytest_gpu.log; tail -5 gpurun_out/bench256.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench256.json').read().strip().splitlines()[-1])
print("value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e'], "launches", d['gpu_launches'])
print("roofline", d['roofline'])
for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s  share {v['share']:.3f}  launches/step {v['launches_per_step']}")
print("cpu", d['cpu_baseline'])
PY
