#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_batches_solve_ws -s 6 -c 1 -o gpurun_out/r2_prof_ws32 python scripts/solve_microbench.py 32 256 > gpurun_out/r2_ncu_ws32.log 2>&1
tail -2 gpurun_out/r2_ncu_ws32.log
python scripts/solve_microbench.py 32 256 2>/dev/null | tail -1
python scripts/solve_microbench.py 256 256 2>/dev/null | tail -1
