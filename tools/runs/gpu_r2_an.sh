#!/bin/bash
# Round 2, call AN: parity suite + the driver's N = 1 bench after K2c requests its xi vectors up front and K2 / K2c / K3 launch with PDL.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_an.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_an.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_an.json 2> gpurun_out/bench_r02_an.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02_an.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_an.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
for k, v in d["aux"].get("elementwise", {}).items():
    print(k, round(v["us_per_launch"], 2), round(v["roofline"]["frac"], 3), v["rotating_sets"], json.dumps(v.get("multi_stream"))[:200])
print("dit", json.dumps(d["aux"].get("dit_train", {}).get("bf16"))[:300])
print("sampler", json.dumps(d["aux"].get("sampler"))[:400])
PY
