#!/bin/bash
# Round 2, call C: single-wave kernel after the weight-first / pair-major / pipelined-finalize changes.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_energy.py -m gpu -x -q -k "wave or variants or headline or golden" > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_c.log
tail -3 gpurun_out/pytest_gpu_c.log
{
timeout 300 python tools/trace_energy.py --tune "energy.variant=5"
timeout 300 python tools/trace_energy.py --tune "energy.variant=5,energy.ksmem=1"
timeout 300 python tools/trace_energy.py --dtype bf16 --tune "energy.variant=5,energy.threads=384,energy.nv=1"
} > gpurun_out/trace_c.log 2>&1
cat gpurun_out/trace_c.log
{
echo "== f32 single stream (ksmem: 1 col-major regs, 2 col-major smem, 3/0 pair-major)"
timeout 900 python tools/sweep_energy.py --streams 1 --configs "variant=5;variant=5,ksmem=1;variant=5,ksmem=2;variant=5,threads=384,nv=2;variant=5,threads=384,nv=2,ksmem=2;variant=5,nostore=1;variant=3"
echo "== f32 nograd"
timeout 600 python tools/sweep_energy.py --streams 1 --nograd --configs "variant=5"
echo "== bf16 single stream"
timeout 900 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=5,threads=384,nv=1;variant=5,threads=384,nv=1,ksmem=1;variant=5,threads=384,nv=1,ksmem=2;variant=5,threads=256,nv=2;variant=3"
echo "== f32 6 streams"
timeout 600 python tools/sweep_energy.py --streams 6 --configs "variant=5;variant=3"
} > gpurun_out/sweep_c.log 2>&1
cat gpurun_out/sweep_c.log
