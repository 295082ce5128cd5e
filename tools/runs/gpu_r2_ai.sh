#!/bin/bash
# Round 2, call AI: cache operator of the gradient stores (experiment build): 0 = L1::no_allocate (default), 4 .wb, 5 .cg, 6 .cs, 7 .wt
mkdir -p gpurun_out
timeout 300 python tools/sweep_energy.py --streams 1 --configs "variant=3;variant=3,sthint=4;variant=3,sthint=5;variant=3,sthint=6;variant=3,sthint=7;variant=3"
timeout 300 python tools/sweep_energy.py --streams 6 --configs "variant=3;variant=3,sthint=6;variant=3,sthint=7"
timeout 600 python -m pytest tests/test_gpu_energy.py -m gpu -q -x -k "workspace or finish" 2>&1 | tail -2
