#!/bin/bash
# One gpurun call (1 GPU): ncu captures of the dominant kernels + the launch list of the bench command.
# Every ncu command is preceded by the same command run plainly (exit 0 required).
mkdir -p gpurun_out
P="python tools/profile_energy.py"
set -x
$P > gpurun_out/plain_k1_f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:energy_fused_smem -s 4 -c 3 -f -o gpurun_out/prof_k1_f32 $P > gpurun_out/ncu_k1_f32.log 2>&1
$P --dtype bf16 > gpurun_out/plain_k1_bf16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:energy_fused_smem -s 4 -c 3 -f -o gpurun_out/prof_k1_bf16 $P --dtype bf16 > gpurun_out/ncu_k1_bf16.log 2>&1
$P --m 32 --iters 6 > gpurun_out/plain_k1_m32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:energy_fused_blk -s 2 -c 2 -f -o gpurun_out/prof_k1_m32 $P --m 32 --iters 6 > gpurun_out/ncu_k1_m32.log 2>&1
B="python bench.py --steps 200 --warmup 3 --cpu-seconds 0 --dit-steps 0 --sampler-samples 0 --e2e-steps 3"
$B > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_*.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench.csv
