#!/bin/bash
for cfg in "32 256" "128 256" "256 256"; do
for v in "DB_SOLVE_MINB=1" "DB_SOLVE_MINB=2" "DB_SOLVE_WS=0 DB_SOLVE_MINB=2" "DB_SOLVE_WS=0 DB_SOLVE_MINB=3"; do
  env $v timeout 300 python scripts/solve_microbench.py $cfg 2>/dev/null | tail -1
done
done
