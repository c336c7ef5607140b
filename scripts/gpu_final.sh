#!/bin/bash
# Round-end style validation on one GPU: tests, smoke, the bench line (with CPU baseline and e2e), the reference arm,
# the ncu launch list of the bench command and --set full captures of the dominant kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench256.json 2> gpurun_out/bench256.err; echo "bench rc=$?"
tail -2 gpurun_out/bench256.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench256.json').read().strip().splitlines()[-1])
print("value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e']['value'], "launches", d['gpu_launches'], "cpu", d['cpu_baseline'], "clocks", d['clocks'])
print("roofline", d['roofline'])
for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s  share {v['share']:.3f}  launches/step {v['launches_per_step']}")
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; tail -1 gpurun_out/bench_reference.json | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --gpus 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
cap() { # kernel-regex skip name
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_$3 -f python bench.py --gpus 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$3.log 2>&1
}
cap k_rbwd_regs 47 xbwd
cap k_batches_solve_flat 2 solve
cap k_pointwise_pairs 2 pointwise
cap k_batches_matvec 2 matvec
cap k_chbwd_regs 4 zbwd
cap k_chfwd_regs 4 zfwd
ls -la gpurun_out/*.ncu-rep | tail -8
