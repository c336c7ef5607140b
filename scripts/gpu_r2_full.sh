#!/bin/bash
# GPU suite + full default bench (parity gate, cpu baseline) + reference arm + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.log; tail -4 gpurun_out/r2f_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "rc=$?"; tail -2 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print("value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e']['value'], "launches", d['gpu_launches'], "parity", d['parity']['ok'], d['parity']['max_rel'], "cpu", d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s  share {v['share']:.3f}")
print(d['roofline'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2f_launches.log 2>&1
wc -l gpurun_out/r2f_launches.csv
