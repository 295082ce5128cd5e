#!/bin/bash
# Round 2, call D: cp.async loader of the TMA-staged kernel (variant 3) against its TMA loader, single launch.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_energy.py -m gpu -x -q -k "variants or ragged or sweep" > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_d.log
tail -3 gpurun_out/pytest_gpu_d.log
{
timeout 300 python tools/trace_energy.py --tune "energy.variant=3,energy.loader=2"
timeout 300 python tools/trace_energy.py --tune "energy.variant=3,energy.loader=2,energy.nv=1"
timeout 300 python tools/trace_energy.py --dtype bf16 --tune "energy.variant=3,energy.loader=2"
} > gpurun_out/trace_d.log 2>&1
cat gpurun_out/trace_d.log
{
echo "== f32 single stream"
timeout 900 python tools/sweep_energy.py --streams 1 --configs "variant=3,loader=1;variant=3,loader=2;variant=3,loader=2,nv=1;variant=3,loader=2,nv=3;variant=3,loader=2,cluster=2;variant=5"
echo "== f32 nograd"
timeout 600 python tools/sweep_energy.py --streams 1 --nograd --configs "variant=3,loader=2;variant=3,loader=1"
echo "== bf16 single stream"
timeout 900 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3,loader=1;variant=3,loader=2;variant=3,loader=2,nv=1"
echo "== 6 streams"
timeout 600 python tools/sweep_energy.py --streams 6 --configs "variant=3,loader=1;variant=3,loader=2"
timeout 600 python tools/sweep_energy.py --streams 6 --dtype bf16 --configs "variant=3,loader=1;variant=3,loader=2"
} > gpurun_out/sweep_d.log 2>&1
cat gpurun_out/sweep_d.log
