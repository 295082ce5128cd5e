#!/bin/bash
# Round 2, call K: tensor-core energy kernel (variant 7) bring-up.
mkdir -p gpurun_out
{
timeout 120 python tools/check_tc.py --m 32 --D 256 --B 3; echo "rc=$?"
timeout 120 python tools/check_tc.py --m 32 --D 3072 --B 8; echo "rc=$?"
timeout 120 python tools/check_tc.py --m 16 --D 3072 --B 150; echo "rc=$?"
timeout 120 python tools/check_tc.py --m 32 --D 12288 --B 4; echo "rc=$?"
} > gpurun_out/check_tc.log 2>&1
cat gpurun_out/check_tc.log
{
echo "== bf16 m=32 D=3072 single stream: tc vs blocked"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --m 32 --configs "variant=7;variant=4"
echo "== bf16 m=16"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --m 16 --configs "variant=7;variant=4"
echo "== 4 streams"
timeout 300 python tools/sweep_energy.py --streams 4 --dtype bf16 --m 32 --configs "variant=7;variant=4"
timeout 300 python tools/sweep_energy.py --streams 4 --dtype bf16 --m 16 --configs "variant=7;variant=4"
} > gpurun_out/sweep_k.log 2>&1
cat gpurun_out/sweep_k.log
