#!/bin/bash
# Round 2, call P: tensor-core kernel: what paces the MMAs?
mkdir -p gpurun_out
for h in 0 2 16 4 6; do
  echo "== ldhint=$h (2 no Gram MMAs, 4 no gradient MMAs, 16 Gram B operand from another slot)"
  timeout 300 python tools/trace_energy.py --dtype bf16 --m 32 --tune "energy.variant=7,energy.ldhint=$h" | grep -A12 "tensor-core kernel stamps\|us/launch" | grep "us/launch\|gram_committed\|conf_pass_done\|coef_ready\|epilogue_done\|row_done\|pass2_mmas"
done > gpurun_out/trace_p.log 2>&1
cat gpurun_out/trace_p.log
