#!/bin/bash
# First GPU session after round 2: everything that was finished after the round's GPU budget was spent (profiles/README.md).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu_next.sh'
mkdir -p gpurun_out
python -m pytest tests/test_gpu_t2_sphere.py tests/test_gpu_t3_complex.py tests/test_gpu_t6_shell.py tests/test_gpu_t7_sphere_lbvp.py tests/test_gpu_t8_expressions.py \
    tests/test_gpu_t9_plugins.py -q > gpurun_out/next_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/next_pytest.log
for mode in 0 3; do
  DB_BANDED_MODE=$mode python scripts/sphere_bench.py --steps 20 > gpurun_out/next_sphere_mode$mode.json 2> gpurun_out/next_sphere_mode$mode.err
done
python scripts/sphere_bench.py --problem shell_convection --nphi 128 --ntheta 64 --nr 64 --steps 10 > gpurun_out/next_shell_128x64x64.json 2> gpurun_out/next_shell.err
tail -3 gpurun_out/next_pytest.log; cat gpurun_out/next_sphere_mode3.json | head -c 600
