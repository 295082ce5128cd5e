#!/bin/bash
# Round 2, call Q: full GPU suite + config-3 sweep (all betas) with the tensor-core kernel.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_q.log
tail -8 gpurun_out/pytest_gpu_q.log
timeout 900 python tools/sweep_cfg3.py --dtypes bf16 --ms 16,32 --Ds 3072,12288 > gpurun_out/sweep_cfg3_bf16_tc.jsonl 2> gpurun_out/sweep_cfg3.err; echo "sweep rc=$?"
cat gpurun_out/sweep_cfg3_bf16_tc.jsonl
timeout 600 python tools/sweep_cfg3.py --dtypes bf16 --ms 16,32 --Ds 3072 --betas 0.1 --tune energy.variant=4 > gpurun_out/sweep_cfg3_bf16_blk.jsonl 2>> gpurun_out/sweep_cfg3.err
cat gpurun_out/sweep_cfg3_bf16_blk.jsonl
