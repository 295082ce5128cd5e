#!/bin/bash
# Round 2, call AW: K2c after the instruction trim (unguarded block of 8 draws, 32-bit token index): parity, then timing.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_step.py -m gpu -q -x -k "concat or fused_io or trainer_cuda_graph" > gpurun_out/pytest_gpu_aw.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/pytest_gpu_aw.log
timeout 60 python tools/time_elementwise.py --only K2c 2>&1 | grep -v Warning | tee gpurun_out/k2c_after_trim.log
