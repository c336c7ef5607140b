#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python scripts/fft_microbench.py 256 2>&1 | tee gpurun_out/fft_microbench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fft -s 20 -c 3 -o gpurun_out/prof_fft python scripts/fft_microbench.py 256 > gpurun_out/ncu_fft.log 2>&1
ls -la gpurun_out/*.ncu-rep
