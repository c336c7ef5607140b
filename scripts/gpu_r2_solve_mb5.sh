#!/bin/bash
for cfg in "32 256" "256 256"; do
for v in "DB_SOLVE_CPS=1" "DB_SOLVE_CPS=2" "DB_SOLVE_CPS=4" "DB_SOLVE_CPS=4 DB_SOLVE_STAGES=3" "DB_SOLVE_CPS=8 DB_SOLVE_STAGES=2" "DB_SOLVE_CPS=4 DB_SOLVE_PF=50"; do
  env $v timeout 300 python scripts/solve_microbench.py $cfg 2>/dev/null | tail -1
done
done
