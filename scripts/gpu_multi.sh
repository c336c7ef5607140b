#!/bin/bash
# usage: gpu_multi.sh "2 4 8"   (run under gpurun --gpus max)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_solver.py -m gpu -q -k two_gpu > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log; tail -3 gpurun_out/pytest_2gpu.log
for n in $1; do
  if [ "$n" = "1" ]; then
    timeout 1200 python bench.py --gpus 1 --size 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
  else
    NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --size 256 --steps 10 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "N=$n rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/scale_$n.err | tail -3
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/scale_$n.json').read().strip().splitlines()[-1])
    print("N=$n value", round(d['value'],2), "ms/step", round(d['ms_per_step'],2), "e2e", round(d['e2e']['value'],2) if d.get('e2e') else None)
    for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s  share {v['share']:.3f}")
except Exception as e:
    print("no result", e)
PY
done
