#!/bin/bash
# Round 2, call L: tensor-core energy kernel: mixed entry, timeline, tau = 2^-8.
mkdir -p gpurun_out
{
timeout 120 python tools/check_tc.py --m 32 --D 3072 --B 8; echo "rc=$?"
timeout 120 python tools/check_tc.py --m 16 --D 3072 --B 150 | head -8; echo "rc=$?"
timeout 120 python tools/check_tc.py --m 32 --D 12288 --B 4 | head -8; echo "rc=$?"
} > gpurun_out/check_tc.log 2>&1
cat gpurun_out/check_tc.log
{
timeout 300 python tools/trace_energy.py --dtype bf16 --m 32 --tune "energy.variant=7"
timeout 300 python tools/trace_energy.py --dtype bf16 --m 16 --tune "energy.variant=7"
timeout 300 python tools/trace_energy.py --dtype bf16 --m 32 --B 4 --tune "energy.variant=7"
} > gpurun_out/trace_l.log 2>&1
cat gpurun_out/trace_l.log
