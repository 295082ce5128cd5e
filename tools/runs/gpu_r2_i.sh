#!/bin/bash
# Round 2, call I: row-pipelined cluster kernel (variant 6): parity, single-stream sweep, timeline; copy speed-of-light.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_energy.py -m gpu -x -q -k "pipe" > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_i.log
tail -15 gpurun_out/pytest_gpu_i.log
{
echo "== f32 single stream"
timeout 900 python tools/sweep_energy.py --streams 1 --configs "variant=3;variant=6,cluster=2;variant=6,cluster=4;variant=6,cluster=8;variant=6,cluster=4,threads=256;variant=6,cluster=8,threads=256;variant=6,cluster=2,threads=256;variant=6,cluster=4,threads=96,cols=4;variant=6,cluster=4,threads=192,cols=4;variant=6,cluster=8,threads=96,cols=4;variant=6,cluster=4,window=2;variant=6,cluster=4,window=3;variant=6,cluster=8,window=3;variant=6,cluster=8,window=5;variant=6,cluster=8,window=6;variant=6,cluster=8,window=5,threads=256"
echo "== bf16 single stream"
timeout 900 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3;variant=6,cluster=2;variant=6,cluster=4;variant=6,cluster=8;variant=6,cluster=4,threads=256;variant=6,cluster=8,threads=256"
echo "== f32 6 streams"
timeout 600 python tools/sweep_energy.py --streams 6 --configs "variant=3;variant=6,cluster=4;variant=6,cluster=8"
} > gpurun_out/sweep_i.log 2>&1
cat gpurun_out/sweep_i.log
{
timeout 300 python tools/trace_energy.py --tune "energy.variant=6,energy.cluster=4"
timeout 300 python tools/trace_energy.py --tune "energy.variant=6,energy.cluster=8"
timeout 300 python tools/trace_energy.py --tune "energy.variant=6,energy.cluster=4,energy.threads=256"
} > gpurun_out/trace_i.log 2>&1
cat gpurun_out/trace_i.log
timeout 600 tools/ubench/copy_sol > gpurun_out/copy_sol.log 2>&1; echo "copy_sol rc=$?"
cat gpurun_out/copy_sol.log
