#!/bin/bash
# Round 2, call J: Philox-fused K3 (bit-identity with torch.randn_like), sampler paths, pipe kernel mixed entry; copy SOL sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_energy.py -m gpu -x -q -k "philox or sampler or pipe" > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_j.log
tail -30 gpurun_out/pytest_gpu_j.log
timeout 600 tools/ubench/copy_sol > gpurun_out/copy_sol.log 2>&1; echo "copy_sol rc=$?"
cat gpurun_out/copy_sol.log
