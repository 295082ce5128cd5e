#!/bin/bash
# Round 2, call Y: tensor-core kernel: depth of the TMA ring.
for m in 32 16; do
  echo "== m=$m one stream"; timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --m $m --configs "variant=7,nv=3;variant=7,nv=4;variant=7,nv=5;variant=7,nv=6"
  echo "== m=$m four streams"; timeout 300 python tools/sweep_energy.py --streams 4 --dtype bf16 --m $m --configs "variant=7,nv=3;variant=7,nv=4;variant=7,nv=5;variant=7,nv=6"
done 2>&1 | cut -c1-250
timeout 200 python tools/check_tc.py --m 32 --D 3072 --B 8 | head -5
