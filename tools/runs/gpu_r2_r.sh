#!/bin/bash
# Round 2, call R: full GPU suite after the tensor-core kernel / fused-noise sampler changes.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_r.log
tail -12 gpurun_out/pytest_gpu_r.log
