#!/bin/bash
# Round 2, call AT: parity suite after the B-aware cluster choice of the TMA-staged kernel, then the auto plan at B = 16..128.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_at.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_at.log
{
for dt in f32 bf16; do for B in 16 32 64 96 128; do
  echo "== $dt B=$B one stream, auto plan"
  timeout 120 python tools/sweep_energy.py --streams 1 --dtype $dt --B $B --configs "variant=0"
done; done
echo "== f32 B=32 / 64, four streams: auto vs whole rows"
timeout 120 python tools/sweep_energy.py --streams 4 --B 32 --configs "variant=0;variant=3,cluster=1"
timeout 120 python tools/sweep_energy.py --streams 4 --B 64 --configs "variant=0;variant=3,cluster=1"
} 2>&1 | grep -v Warning | tee gpurun_out/k1_small_batch_auto.log
