#!/bin/bash
# merged sign-equivalent components: GPU suite + bench + ncu of solve / matvec / move
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_gpu.log; tail -5 gpurun_out/r2m_pytest_gpu.log
timeout 900 python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench256.json 2> gpurun_out/r2m_bench256.err; echo "rc=$?" >> gpurun_out/r2m_bench256.err
tail -3 gpurun_out/r2m_bench256.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m_bench256.json').read().strip().splitlines()[-1])
print("value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e']['value'], "launches", d['gpu_launches'], "parity", d['parity']['ok'], d['parity']['max_rel'])
for k, v in d['kernels'].items(): print(f"  {k:24s} {v['ms_per_step']:8.2f} ms/step  {v['gbps']:8.1f} GB/s  share {v['share']:.3f}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batches_solve -s 4 -c 1 -o gpurun_out/r2m_prof_solve python bench.py --size 256 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2m_ncu_solve.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
for P in 8 2; do
DB_CHEB_P=$P timeout 600 python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2m_bench_chP$P.json 2> gpurun_out/r2m_bench_chP$P.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r2m_bench_chP$P.json').read().strip().splitlines()[-1])
print("CHEB_P=$P ms/step", d['ms_per_step'], {k: (round(v['ms_per_step'],2), round(v['gbps'])) for k, v in d['kernels'].items() if 'axis2' in k})
PY
done
