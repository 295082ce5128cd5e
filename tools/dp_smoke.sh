#!/bin/bash
# Multi-GPU smoke on one box: bench.py at N ranks and a short data-parallel training run.
# usage (GPU box): bash tools/dp_smoke.sh N
set -e
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20000 --warmup 50 2>gpurun_out/bench_n$N.err | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    -m ddm_b200.launcher --synthetic --epochs 2 --steps-per-epoch 30 --log-every 10 --out gpurun_out/dp_run_n$N \
    --sample-batch 16 --sample-steps 5 2>gpurun_out/launcher_n$N.err | tail -8
ls gpurun_out/dp_run_n$N
