#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
DB_SOLVE_STAGES=2 timeout 1500 python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench256.json 2> gpurun_out/bench256.err; echo "rc=$?" >> gpurun_out/bench256.err
tail -3 gpurun_out/bench256.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench256.json').read().strip().splitlines()[-1])
print("value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e']['value'], "launches", d['gpu_launches'])
for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s  share {v['share']:.3f}")
PY
for k in k_pointwise_v3 k_batches_matvec k_batches_solve_flat "k_batches_move"; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k python bench.py --size 256 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
