#!/bin/bash
# Round 2, call AF: with the polled finish in place, D-split clusters and the 8-warp build on ONE launch (bf16 is bound by
# its arithmetic on 4 warps per SM; fp32 for completeness); blocked kernel with the polled finish.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_energy.py -m gpu -q -x -k "finish or blocked or headline" > gpurun_out/pytest_gpu_af.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_af.log
echo "== bf16 one stream"
timeout 200 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3;variant=3,cluster=2;variant=3,cluster=4;variant=3,threads=256,nv=1;variant=3,cluster=2,threads=64"
echo "== bf16 six streams"
timeout 200 python tools/sweep_energy.py --streams 6 --dtype bf16 --configs "variant=3;variant=3,cluster=2;variant=3,cluster=4"
echo "== f32 one stream"
timeout 200 python tools/sweep_energy.py --streams 1 --configs "variant=3;variant=3,cluster=2;variant=3,cluster=4;variant=3,threads=256,nv=1"
echo "== f32 six streams"
timeout 200 python tools/sweep_energy.py --streams 6 --configs "variant=3;variant=3,cluster=2"
echo "== m=16 / m=32 f32 blocked kernel one stream: ticket vs polled"
timeout 200 python tools/sweep_energy.py --streams 1 --m 16 --configs "variant=4,finish=1;variant=4,finish=2"
timeout 200 python tools/sweep_energy.py --streams 1 --m 32 --configs "variant=4,finish=1;variant=4,finish=2"
