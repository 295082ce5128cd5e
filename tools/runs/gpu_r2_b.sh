#!/bin/bash
# Round 2, call B: diagnostics of the single-wave kernel — no-gradient / no-store timings, L2 eviction hints,
# SM clock inside the kernel, and one ncu capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_energy.py -m gpu -x -q -k "wave or variants or headline" > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_b.log
tail -3 gpurun_out/pytest_gpu_b.log
{
timeout 300 python tools/trace_energy.py --tune "energy.variant=5"
timeout 300 python tools/trace_energy.py --tune "energy.variant=5,energy.nostore=1"
} > gpurun_out/trace_b.log 2>&1
cat gpurun_out/trace_b.log
{
echo "== nograd (forward only: no pass 2, no stores)"
timeout 600 python tools/sweep_energy.py --streams 1 --nograd --configs "variant=5;variant=3"
echo "== nostore (pass 2 computed, nothing stored)"
timeout 600 python tools/sweep_energy.py --streams 1 --configs "variant=5,nostore=1"
echo "== L2 hints (ld,st): 0 normal 1 evict_first 2 evict_last 3 unchanged"
timeout 900 python tools/sweep_energy.py --streams 1 --configs "variant=5;variant=5,ldhint=1;variant=5,sthint=1;variant=5,ldhint=1,sthint=1;variant=5,ldhint=1,sthint=2;variant=5,ldhint=2,sthint=1;variant=5,ldhint=3,sthint=3;variant=5,ldhint=0,sthint=2;variant=5,ldhint=2,sthint=0"
} > gpurun_out/sweep_b.log 2>&1
cat gpurun_out/sweep_b.log
P="python tools/profile_energy.py"
$P > gpurun_out/plain_k1_f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:energy_fused_wave -s 4 -c 2 -f -o gpurun_out/prof_k1_wave_f32 $P > gpurun_out/ncu_k1_wave_f32.log 2>&1
tail -3 gpurun_out/ncu_k1_wave_f32.log
