#!/bin/bash
# Round 2, call O: tensor-core energy kernel v3 (single centred bf16 tile, x0 staged, TMA-store epilogue).
mkdir -p gpurun_out
{
timeout 120 python tools/check_tc.py --m 32 --D 256 --B 3 | head -14; echo "rc=$?"
timeout 120 python tools/check_tc.py --m 32 --D 3072 --B 8; echo "rc=$?"
timeout 120 python tools/check_tc.py --m 16 --D 3072 --B 150; echo "rc=$?"
timeout 120 python tools/check_tc.py --m 32 --D 12288 --B 4 | head -14; echo "rc=$?"
} > gpurun_out/check_tc.log 2>&1
cat gpurun_out/check_tc.log
{
timeout 300 python tools/trace_energy.py --dtype bf16 --m 32 --tune "energy.variant=7"
timeout 300 python tools/trace_energy.py --dtype bf16 --m 16 --tune "energy.variant=7"
} > gpurun_out/trace_o.log 2>&1
grep -A14 "tensor-core kernel stamps" gpurun_out/trace_o.log; grep "us/launch" gpurun_out/trace_o.log
{
echo "== bf16 m=32 D=3072 single stream: tc vs blocked"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --m 32 --configs "variant=7;variant=4"
echo "== bf16 m=16"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --m 16 --configs "variant=7;variant=4"
echo "== 4 streams"
timeout 300 python tools/sweep_energy.py --streams 4 --dtype bf16 --m 32 --configs "variant=7;variant=4"
timeout 300 python tools/sweep_energy.py --streams 4 --dtype bf16 --m 16 --configs "variant=7;variant=4"
} > gpurun_out/sweep_o.log 2>&1
cat gpurun_out/sweep_o.log
