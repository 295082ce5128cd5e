#!/bin/bash
# Round 2, call AQ (2 GPUs): the two-device test of the per-device host caches, then the driver's N = 2 bench on the final code.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step.py -m gpu -q -k second_device -rs > gpurun_out/pytest_gpu_aq.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_aq.log
bash tools/runs/gpu_r2_w.sh 2
