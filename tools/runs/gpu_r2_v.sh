#!/bin/bash
# Round 2, call V: centred forms in the blocked kernel (m = 16 / 24 / 32): parity + timing.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_energy.py -m gpu -x -q > gpurun_out/pytest_gpu_v.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_v.log
tail -6 gpurun_out/pytest_gpu_v.log
timeout 900 python tools/sweep_cfg3.py --dtypes f32,bf16 --ms 16,32 --Ds 3072,12288 --betas 0.1 --tune energy.variant=4 > gpurun_out/sweep_cfg3_blk_centred.jsonl 2> gpurun_out/sweep_v.err; echo "sweep rc=$?"
cut -c1-330 gpurun_out/sweep_cfg3_blk_centred.jsonl
