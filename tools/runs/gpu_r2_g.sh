#!/bin/bash
# Round 2, call G: first-data latency of a single launch against the number of rows (latency or bandwidth?).
mkdir -p gpurun_out
{
for B in 4 16 64 128; do
timeout 300 python tools/trace_energy.py --B $B --tune "energy.variant=3,energy.loader=1" | grep -E "kernel:|inputs_ready->first|first_chunk->pass1|period|coef_ready->pass2|entry->inputs"
timeout 300 python tools/trace_energy.py --B $B --tune "energy.variant=3,energy.loader=2,energy.window=8" | grep -E "kernel:|inputs_ready->first|first_chunk->pass1|period|coef_ready->pass2|entry->inputs"
done
} > gpurun_out/trace_g.log 2>&1
cat gpurun_out/trace_g.log
