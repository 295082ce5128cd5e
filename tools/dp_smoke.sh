#!/bin/bash
# Multi-GPU smoke on one box: bench.py at N ranks and a short data-parallel training run (every command bounded).
# usage (GPU box): bash tools/dp_smoke.sh N
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus $N --warmup 50 --aux-timeout 120 2>gpurun_out/bench_n$N.err | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-400
echo "bench rc=${PIPESTATUS[0]}"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
    print("aux", json.dumps(d.get("aux"))[:900])
except Exception as e:
    print("no bench line:", e)
PY
timeout 240 $TR --master-port 29512 -m ddm_b200.launcher --synthetic --epochs 2 --steps-per-epoch 30 --log-every 10 \
    --out /tmp/dp_run_n$N --sample-batch 16 --sample-steps 5 2>gpurun_out/launcher_n$N.err | tail -6
echo "launcher rc=${PIPESTATUS[0]}"
ls /tmp/dp_run_n$N; cp /tmp/dp_run_n$N/train_history.json gpurun_out/train_history_n$N.json
