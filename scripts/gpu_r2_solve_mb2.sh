#!/bin/bash
mkdir -p gpurun_out
for cfg in "32 256" "128 256" "256 256"; do
  for v in "DB_SOLVE_WS=0" "DB_SOLVE_WS=1" "DB_SOLVE_WS=1 DB_SOLVE_STAGES=4" "DB_SOLVE_WS=1 DB_SOLVE_STAGES=12" "DB_SOLVE_WS=1 DB_SOLVE_MINB=4"; do
    env $v timeout 300 python scripts/solve_microbench.py $cfg 2>gpurun_out/mb.err | tail -1 || tail -5 gpurun_out/mb.err
  done
done
