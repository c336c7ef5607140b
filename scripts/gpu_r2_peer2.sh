#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py -m gpu -q -x -k "two_gpu" > gpurun_out/r2_pytest_2gpu_peer.log 2>&1; tail -15 gpurun_out/r2_pytest_2gpu_peer.log
run() { name=$1; n=$2; shift; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/pe_$name.json 2> gpurun_out/pe_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/pe_$name.json').read().strip().splitlines()[-1])
    print("$name steps/s", round(d['value'],2), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],2), "parity", d['parity']['ok'], d['parity']['max_rel'], "chk", d['state_checksum']['sum_sq'])
    print("   ", {k: (round(v['ms_per_step'],2), round(v['gbps'])) for k, v in d['kernels'].items()})
except Exception as e: print("$name failed", e); print(open('gpurun_out/pe_$name.err').read()[-2500:])
PY
}
run peer 2 DB_PEER_TRANSPOSE=1
run nccl 2 DB_PEER_TRANSPOSE=0
grep -i "peer-memory\|warn" gpurun_out/pe_peer.err | head -5
