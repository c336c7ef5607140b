#!/bin/bash
# 8-GPU box: 2-GPU NCCL tests, then bench at N = 8, 4 (kernel breakdown), solve variants at N = 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_solver.py -m gpu -q -x -k "two_gpu" > gpurun_out/r2_pytest_2gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_2gpu.log
run() { # name, ngpu, env...
  name=$1; n=$2; shift; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/sc_$name.json 2> gpurun_out/sc_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/sc_$name.json').read().strip().splitlines()[-1])
    print("$name steps/s", round(d['value'],2), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],2), "parity", d['parity']['ok'], d['parity']['max_rel'], d['parity'].get('blocked_transposes'), "chk", d['state_checksum']['sum_sq'])
    print("   ", {k: (round(v['ms_per_step'],2), round(v['gbps'])) for k, v in d['kernels'].items()})
except Exception as e: print("$name failed", e); print(open('gpurun_out/sc_$name.err').read()[-1500:])
PY
}
run n8 8
run n8_rt4 8 DB_SOLVE_RT=4
run n8_pipe 8 DB_SOLVE_PIPE=1
run n4 4
run n2 2
