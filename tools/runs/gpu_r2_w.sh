#!/bin/bash
# Round 2, call W (N GPUs): the driver's multi-rank bench invocation.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; echo "bench rc=$?"
tail -c 1200 gpurun_out/bench_r02_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_n$N.json").read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "e2e", "clocks")}, indent=1)[:1500])
    a = d.get("aux", {})
    for k in ("k1_with_weight_allreduce", "dp_parity", "dp_parity_bf16", "sampler"):
        print(k, json.dumps(a.get(k))[:900])
    print("dit", json.dumps({k: v for k, v in a.get("dit_train", {}).items() if k in ("bf16", "tf32", "fp32")})[:900])
except Exception as e:
    print("parse failed", e)
PY
timeout 300 $TR --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 3 2>/dev/null | tail -1 | cut -c1-300
