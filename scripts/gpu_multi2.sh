#!/bin/bash
# 2-GPU validation: NCCL tests (incl. blocked transposes vs pack/unpack) + bench at N=2 with both transpose paths
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_solver.py -m gpu -q -k two_gpu > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log; tail -4 gpurun_out/pytest_2gpu.log
for mode in 1 0; do
  DB_BLOCKED_TRANSPOSE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus 2 --size 256 --steps 8 --warmup 3 --no-e2e > gpurun_out/scale2_blocked$mode.json 2> gpurun_out/scale2_blocked$mode.err
  echo "blocked=$mode rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/scale2_blocked$mode.err | tail -3
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/scale2_blocked$mode.json').read().strip().splitlines()[-1])
    print("blocked=$mode value", round(d['value'],2), "ms/step", round(d['ms_per_step'],2))
    for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s")
except Exception as e:
    print("no result", e)
PY
done
