#!/bin/bash
# usage: gpu_r2_512.sh NGPU  -- 512^3 (parity gate on) and 256^3 on NGPU GPUs
n=$1
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $n "$@" > gpurun_out/s5_$name.json 2> gpurun_out/s5_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/s5_$name.json').read().strip().splitlines()[-1])
    print("$name", d['metric'], "steps/s", round(d['value'],3), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],3) if d.get('e2e') else None, "parity", d['parity'] and (d['parity']['ok'], d['parity']['max_rel'], d['parity']['shape']), "chk", d['state_checksum']['sum_sq'], "setup_s", round(d['config']['setup_seconds'],1))
    print("   ", {k: (round(v['ms_per_step'],2), round(v['gbps'])) for k, v in d['kernels'].items()})
except Exception as e: print("$name failed", e); print(open('gpurun_out/s5_$name.err').read()[-2500:])
PY
}
run 512_n$n --size 512 --steps 5 --warmup 3 --no-cpu-baseline
run 256_n$n --size 256 --steps 10 --warmup 3 --no-cpu-baseline
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
