#!/bin/bash
# Round 2, call U: centred pass 2 in the headline kernel: parity + timing.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_gpu_energy.py tests/test_gpu_step.py -m gpu -x -q > gpurun_out/pytest_gpu_u.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_u.log
tail -6 gpurun_out/pytest_gpu_u.log
{
echo "== f32 1 stream / 6 streams"
timeout 300 python tools/sweep_energy.py --streams 1 --configs "variant=3"
timeout 300 python tools/sweep_energy.py --streams 6 --configs "variant=3"
echo "== bf16 1 stream / 6 streams"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3"
timeout 300 python tools/sweep_energy.py --streams 6 --dtype bf16 --configs "variant=3"
timeout 300 python tools/sweep_energy.py --streams 6 --dtype bf16 --configs "variant=3,cols=2;variant=3,ctas=4"
} > gpurun_out/sweep_u.log 2>&1
cat gpurun_out/sweep_u.log
timeout 300 python tools/trace_energy.py | grep "us/launch\|coef_ready->pass2_done\|pass1_done->coef_ready\|first_chunk->pass1"
