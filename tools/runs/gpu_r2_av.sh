#!/bin/bash
# Round 2, call AV: ncu --set full of K2c and K3 (flat form) on the final code; the plain command first.
mkdir -p gpurun_out
P="python tools/time_elementwise.py"
timeout 60 $P --only K2c,K3 > gpurun_out/plain_elementwise.log 2>&1 || exit 1
timeout 100 ncu --set full --clock-control none --import-source on -k regex:forward_marginal_concat -s 6 -c 2 -f -o gpurun_out/prof_k2c $P --only K2c > gpurun_out/ncu_k2c.log 2>&1
timeout 100 ncu --set full --clock-control none --import-source on -k regex:bridge_step_flat -s 6 -c 2 -f -o gpurun_out/prof_k3 $P --only K3 > gpurun_out/ncu_k3.log 2>&1
tail -2 gpurun_out/ncu_k2c.log gpurun_out/ncu_k3.log; ls -la gpurun_out/*.ncu-rep
