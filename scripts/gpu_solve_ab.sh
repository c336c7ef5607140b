#!/bin/bash
mkdir -p gpurun_out
for order in level natural; do
  DB_SOLVE_ORDER=$order timeout 900 python bench.py --size 256 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$order.json 2> gpurun_out/bench_$order.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$order.json').read().strip().splitlines()[-1])
k = d['kernels']
print("$order", "steps/s", round(d['value'],2), "solve ms", round(k['pencil_solve']['ms_per_step'],2), "GB/s", round(k['pencil_solve']['gbps']), "fft bwd0", round(k['transform_bwd_axis0']['ms_per_step'],2), "fwd0", round(k['transform_fwd_axis0']['ms_per_step'],2))
PY
done
