#!/bin/bash
# Round 2, call AJ: column-chunk size of the TMA-staged kernel on one launch, with the polled finish (nv = 16-byte vectors per thread and chunk)
timeout 300 python tools/sweep_energy.py --streams 1 --configs "variant=3;variant=3,nv=1;variant=3,nv=3;variant=3;variant=3,nv=1"
timeout 300 python tools/sweep_energy.py --streams 6 --configs "variant=3;variant=3,nv=1"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3;variant=3,nv=1;variant=3,nv=3"
timeout 300 python tools/sweep_energy.py --streams 6 --dtype bf16 --configs "variant=3;variant=3,nv=1"
