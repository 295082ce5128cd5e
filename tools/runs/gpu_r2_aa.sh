#!/bin/bash
# Round 2, call AA: GPU suite after the cluster change; D = 12288 timing; ncu of the bf16 kernel.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_aa.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_aa.log
timeout 300 python tools/sweep_energy.py --streams 1 --D 12288 --configs "variant=3"
timeout 300 python tools/sweep_energy.py --streams 4 --D 12288 --configs "variant=3"
timeout 300 python tools/sweep_energy.py --streams 4 --D 12288 --dtype bf16 --configs "variant=3"
P="python tools/profile_energy.py"
$P --dtype bf16 > gpurun_out/plain_k1_bf16.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_fused_smem -s 4 -c 3 -f -o gpurun_out/prof_r02_k1_bf16 $P --dtype bf16 > gpurun_out/ncu_k1_bf16.log 2>&1
tail -1 gpurun_out/ncu_k1_bf16.log
