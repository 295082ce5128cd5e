#!/bin/bash
# One gpurun call: GPU parity tests, then a timeline trace and a small sweep of the fused energy kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/trace_energy.py > gpurun_out/trace.log 2>&1
cat gpurun_out/trace.log
CFG="variant=3,pdl=1;variant=3,pdl=1,nv=1;variant=3,pdl=1,nv=3"
{
for s in 1 4; do for dt in f32 bf16; do
  echo "== streams=$s dtype=$dt"; python tools/sweep_energy.py --streams $s --dtype $dt --configs "$CFG"
done; done
} > gpurun_out/sweep3.log 2>&1
cat gpurun_out/sweep3.log
