#!/bin/bash
mkdir -p gpurun_out
for cfg in "32 256" "64 256" "128 256"; do
  for v in "DB_SOLVE_DEEP=0" "DB_SOLVE_DEEP=1 DB_SOLVE_DEEP_MPC=1" "DB_SOLVE_DEEP=1 DB_SOLVE_DEEP_MPC=2" "DB_SOLVE_DEEP=1 DB_SOLVE_DEEP_MPC=4" "DB_SOLVE_DEEP=0 DB_SOLVE_PIPE=1"; do
    env $v timeout 300 python scripts/solve_microbench.py $cfg 2>gpurun_out/mb.err | tail -1 || tail -5 gpurun_out/mb.err
  done
done
env DB_SOLVE_DEEP=1 DB_SOLVE_DEEP_MPC=4 timeout 300 python scripts/solve_microbench.py 256 256 2>gpurun_out/mb.err | tail -1
env DB_SOLVE_DEEP=1 DB_SOLVE_DEEP_MPC=2 timeout 300 python scripts/solve_microbench.py 256 256 2>gpurun_out/mb.err | tail -1
env DB_SOLVE_DEEP=0 timeout 300 python scripts/solve_microbench.py 256 256 2>gpurun_out/mb.err | tail -1
