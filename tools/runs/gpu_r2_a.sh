#!/bin/bash
# Round 2, call A: GPU parity tests, then timeline traces and a plan sweep of the single-wave kernel (variant 5).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
{
for t in "energy.variant=5" "energy.variant=5,energy.ksmem=1" "energy.variant=5,energy.threads=384,energy.nv=2,energy.ksmem=1" "energy.variant=3"; do
  timeout 300 python tools/trace_energy.py --tune "$t"
done
timeout 300 python tools/trace_energy.py --dtype bf16 --tune "energy.variant=5,energy.threads=384,energy.nv=1,energy.ksmem=1"
} > gpurun_out/trace.log 2>&1
cat gpurun_out/trace.log
F32="variant=5;variant=5,ksmem=1;variant=5,threads=384,nv=2;variant=5,threads=384,nv=2,ksmem=1;variant=5,threads=128,nv=3,ksmem=1;variant=5,pdl=0;variant=3"
BF="variant=5,threads=384,nv=1,ksmem=1;variant=5,threads=384,nv=1;variant=5,threads=256,nv=2;variant=5,threads=256,nv=2,ksmem=1;variant=5,threads=128,nv=3,ksmem=1;variant=3"
{
for s in 1 6; do
  echo "== streams=$s dtype=f32"; timeout 600 python tools/sweep_energy.py --streams $s --dtype f32 --configs "$F32"
  echo "== streams=$s dtype=bf16"; timeout 600 python tools/sweep_energy.py --streams $s --dtype bf16 --configs "$BF"
done
} > gpurun_out/sweep_wave.log 2>&1
cat gpurun_out/sweep_wave.log
