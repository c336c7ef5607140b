#!/bin/bash
# usage: gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_solver.py -m gpu -q -k two_gpu > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log; tail -5 gpurun_out/pytest_2gpu.log
for n in 1 $N; do
  if [ "$n" = "1" ]; then
    timeout 1200 python bench.py --gpus 1 --size 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
  else
    NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --size 256 --steps 10 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "N=$n rc=$?"; tail -3 gpurun_out/scale_$n.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/scale_$n.json').read().strip().splitlines()[-1])
    print("N=$n value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e']['value'] if d.get('e2e') else None)
    for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s  share {v['share']:.3f}")
except Exception as e:
    print("no result", e)
PY
done
