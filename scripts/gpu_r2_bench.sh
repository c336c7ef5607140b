#!/bin/bash
# round 2: default bench (parity gate + reference cpu_baseline), the reference arm as the driver calls it
mkdir -p gpurun_out
nproc > gpurun_out/r2_nproc.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err ) 2> gpurun_out/r2_bench_full.time; echo "rc=$?" >> gpurun_out/r2_bench_full.err
tail -5 gpurun_out/r2_bench_full.err; cat gpurun_out/r2_bench_full.time
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2> gpurun_out/r2_bench_ref.time; echo "rc=$?" >> gpurun_out/r2_bench_ref.err
tail -5 gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.time
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_full.json').read().strip().splitlines()[-1])
print("value", d['value'], "e2e", d['e2e']['value'], "parity", d['parity'], "chk", d['state_checksum'], "cpu", d['cpu_baseline'])
r = json.loads(open('gpurun_out/r2_bench_ref.json').read().strip().splitlines()[-1])
print("ref value", r['value'], r['cpu_baseline'])
PY
