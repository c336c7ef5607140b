#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s8_256.json 2> gpurun_out/s8_256.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/s8_256.json').read().strip().splitlines()[-1])
    print("n8", d['metric'], "steps/s", round(d['value'],3), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],3), "parity", (d['parity']['ok'], d['parity']['max_rel'], d['parity']['blocked_transposes']), "chk", d['state_checksum'])
    print("   ", {k: (round(v['ms_per_step'],2), round(v['gbps'])) for k, v in d['kernels'].items()})
except Exception as e: print("n8 failed", e); print(open('gpurun_out/s8_256.err').read()[-2500:])
PY
