#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/ab_$name.json').read().strip().splitlines()[-1])
    print("$name ms/step", round(d['ms_per_step'],2), {k: (round(v['ms_per_step'],2), round(v['gbps'])) for k, v in d['kernels'].items() if 'solve' in k or 'matvec' in k})
except Exception as e: print("$name failed", e)
PY
}
run pipe DB_SOLVE_PIPE=1
run st4 DB_SOLVE_STAGES=4
run st6 DB_SOLVE_STAGES=6
run st12 DB_SOLVE_STAGES=12
run pipe4 DB_SOLVE_PIPE=1 DB_SOLVE_STAGES=4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chfwd_regs -s 6 -c 1 -o gpurun_out/r2_prof_chfwd $B --steps 1 > gpurun_out/r2_ncu_chfwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chbwd_regs -s 20 -c 1 -o gpurun_out/r2_prof_chbwd $B --steps 1 > gpurun_out/r2_ncu_chbwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batches_matvec -s 4 -c 1 -o gpurun_out/r2_prof_matvec $B --steps 1 > gpurun_out/r2_ncu_matvec.log 2>&1
ls -la gpurun_out/r2_prof_*.ncu-rep
