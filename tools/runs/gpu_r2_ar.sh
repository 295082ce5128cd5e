#!/bin/bash
# Round 2, call AR: K1 with and without programmatic dependent launch on the final code (K2 turned out slower with it);
# L2-resident inputs (2 sets) with 4 and 8 compute warps.
mkdir -p gpurun_out
{
echo "== f32 one stream, pdl A/B"
timeout 200 python tools/sweep_energy.py --streams 1 --configs "variant=3,pdl=1;variant=3,pdl=0;variant=3,pdl=1;variant=3,pdl=0"
echo "== bf16 one stream, pdl A/B"
timeout 200 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3,pdl=1;variant=3,pdl=0"
echo "== f32 six streams, pdl A/B"
timeout 200 python tools/sweep_energy.py --streams 6 --configs "variant=3,pdl=1;variant=3,pdl=0"
echo "== L2-resident (2 sets), one stream: 4 vs 8 compute warps, f32 then bf16"
timeout 200 python tools/sweep_energy.py --streams 1 --sets 2 --configs "variant=3;variant=3,threads=256,nv=1;variant=3,pdl=0"
timeout 200 python tools/sweep_energy.py --streams 1 --sets 2 --dtype bf16 --configs "variant=3;variant=3,threads=256,nv=1;variant=3,pdl=0"
} 2>&1 | grep -v Warning | tee gpurun_out/k1_pdl_ab_final.log
