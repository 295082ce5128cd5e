#!/bin/bash
# Round 2, call E: windowed cp.async loader (variant 3), single launch.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_energy.py -m gpu -x -q -k "variants or ragged" > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_e.log
tail -3 gpurun_out/pytest_gpu_e.log
{
timeout 300 python tools/trace_energy.py --tune "energy.variant=3,energy.loader=2,energy.nv=1,energy.window=6"
timeout 300 python tools/trace_energy.py --tune "energy.variant=3,energy.loader=2,energy.nv=2,energy.window=2"
} > gpurun_out/trace_e.log 2>&1
cat gpurun_out/trace_e.log
{
echo "== f32 single stream"
C=""
for nv in 1 2; do for w in 1 2 3 4 6 8; do C="$C;variant=3,loader=2,nv=$nv,window=$w"; done; done
timeout 1200 python tools/sweep_energy.py --streams 1 --configs "variant=3,loader=1$C"
echo "== f32 nograd"
timeout 600 python tools/sweep_energy.py --streams 1 --nograd --configs "variant=3,loader=2,nv=1,window=6;variant=3,loader=2,nv=1,window=4;variant=3,loader=2,nv=2,window=2"
echo "== bf16 single stream"
timeout 900 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3,loader=2,nv=1,window=3;variant=3,loader=2,nv=1,window=2;variant=3,loader=2,nv=2,window=2;variant=3,loader=2,nv=2,window=1"
} > gpurun_out/sweep_e.log 2>&1
cat gpurun_out/sweep_e.log
