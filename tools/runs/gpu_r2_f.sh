#!/bin/bash
# Round 2, call F: is the ~2.7 us first-data latency of a single launch a TLB effect?  Same kernel, fewer rotating sets.
mkdir -p gpurun_out
{
for n in 40 12 6 2; do
timeout 300 python tools/trace_energy.py --launches $n --tune "energy.variant=3,energy.loader=1" | grep -E "kernel:|inputs_ready->first|first_chunk->pass1|period|coef_ready->pass2"
done
for n in 40 6; do
timeout 300 python tools/trace_energy.py --launches $n --tune "energy.variant=5" | grep -E "kernel:|inputs_ready->first|first_chunk->pass1|period|coef_ready->pass2"
done
} > gpurun_out/trace_f.log 2>&1
cat gpurun_out/trace_f.log
{
for n in 0 20 10 6 5; do echo "== sets=$n"; timeout 600 python tools/sweep_energy.py --streams 1 --sets $n --configs "variant=3,loader=1;variant=5"; done
} > gpurun_out/sweep_f.log 2>&1
cat gpurun_out/sweep_f.log
