#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
echo "regs=1"; timeout 600 python scripts/fft_microbench.py 256 2>&1 | tee gpurun_out/fft_microbench.log
echo "regs=0"; DB_FFT_REGS=0 timeout 600 python scripts/fft_microbench.py 256 2>&1 | tee gpurun_out/fft_microbench_generic.log
timeout 1500 python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench256.json 2> gpurun_out/bench256.err; echo "rc=$?" >> gpurun_out/bench256.err
tail -3 gpurun_out/bench256.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench256.json').read().strip().splitlines()[-1])
print("value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e']['value'], "launches", d['gpu_launches'])
for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s  share {v['share']:.3f}")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rbwd_regs -s 2 -c 1 -o gpurun_out/prof_rbwd_regs python scripts/fft_microbench.py 256 > gpurun_out/ncu_fft.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rfwd_regs -s 2 -c 1 -o gpurun_out/prof_rfwd_regs python scripts/fft_microbench.py 256 >> gpurun_out/ncu_fft.log 2>&1
ls -la gpurun_out/*.ncu-rep
