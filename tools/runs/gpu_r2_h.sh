#!/bin/bash
# Round 2, call H: full GPU test suite + the restructured bench line (N = 1).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_h.log
tail -5 gpurun_out/pytest_gpu_h.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_h.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_h.json").read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "roofline", "e2e", "clocks")}, indent=1)[:3000])
    print(json.dumps(d.get("aux"), indent=1)[:6000])
except Exception as e:
    print("parse failed", e)
PY
