#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --size 128 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench128.json 2> gpurun_out/bench128.err; echo "rc=$?" >> gpurun_out/bench128.err
timeout 1500 python bench.py --size 256 --steps 10 --warmup 3 > gpurun_out/bench256.json 2> gpurun_out/bench256.err; echo "rc=$?" >> gpurun_out/bench256.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > gpurun_out/mem_after.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches128.csv python bench.py --size 128 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench128.log 2>&1
tail -3 gpurun_out/smoke.log; cat gpurun_out/bench128.err | tail -5; cut -c1-1500 gpurun_out/bench128.json; echo; tail -5 gpurun_out/bench256.err; cut -c1-3000 gpurun_out/bench256.json
