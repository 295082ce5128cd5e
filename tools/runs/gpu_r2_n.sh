#!/bin/bash
# Round 2, call N: where does the tensor-core kernel's time go?  (diagnostic knobs via energy.ldhint)
mkdir -p gpurun_out
for h in 0 1 3 7 15 4 8 12; do
  echo "== ldhint=$h (1 no transform, 2 no Gram MMAs, 4 no gradient MMAs, 8 no stores)"
  timeout 300 python tools/trace_energy.py --dtype bf16 --m 32 --tune "energy.variant=7,energy.ldhint=$h" | grep -A12 "tensor-core kernel stamps\|us/launch" | grep -v "^--"
done > gpurun_out/trace_n.log 2>&1
cat gpurun_out/trace_n.log
