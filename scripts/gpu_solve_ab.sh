#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
for ch in 8 16; do
  DB_SOLVE_CH=$ch timeout 900 python bench.py --size 256 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ch$ch.json 2> gpurun_out/bench_ch$ch.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_ch$ch.json').read().strip().splitlines()[-1])
k = d['kernels']
print("CH=$ch", "steps/s", round(d['value'],2), "solve ms", round(k['pencil_solve']['ms_per_step'],2), "GB/s", round(k['pencil_solve']['gbps']), "| transforms ms", round(sum(v['ms_per_step'] for n,v in k.items() if n.startswith('transform')),2))
for n, v in k.items(): print(f"    {n:24s} {v['ms_per_step']:7.2f} ms  {v['gbps']:8.1f} GB/s")
PY
done
timeout 600 python scripts/fft_microbench.py 256 2>&1 | tee gpurun_out/fft_microbench.log
