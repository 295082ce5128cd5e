#!/bin/bash
# Round 2, call AM: the driver's N = 1 bench command after the elementwise harness keeps every launch's outputs distinct.
mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_am.json 2> gpurun_out/bench_r02_am.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02_am.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_am.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
for k, v in d["aux"].get("elementwise", {}).items():
    print(k, round(v["us_per_launch"], 2), round(v["roofline"]["frac"], 3), v["rotating_sets"], json.dumps(v.get("multi_stream"))[:200])
PY
