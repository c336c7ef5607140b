#!/bin/bash
for v in "DB_SOLVE_PF=2" "DB_SOLVE_PF=18" "DB_SOLVE_PF=34" "DB_SOLVE_PF=50"; do
  env $v timeout 300 python scripts/solve_microbench.py 32 256 2>/dev/null | tail -1
done
