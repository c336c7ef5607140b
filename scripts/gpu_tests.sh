#!/bin/bash
# First GPU contact: environment, tests, and timing of the raw kernels.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
python -c "import torch; print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))" > gpurun_out/dev.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
