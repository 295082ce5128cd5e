#!/bin/bash
# Round-end evidence pass on one B200: parity suite, bench lines (ours fp32 / bf16, reference arm), config-3 sweep,
# in-kernel timeline.  tools/gpu_profile.sh holds the ncu captures.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"
timeout 300 python bench.py --dtype bf16 --cpu-seconds 0 --dit-steps 0 --sampler-samples 0 --mmd-samples 0 > gpurun_out/bench_r01_bf16.json 2>> gpurun_out/bench_r01.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench_r01.err
timeout 400 python tools/sweep_cfg3.py > gpurun_out/sweep_cfg3.jsonl 2>&1
{ python tools/trace_energy.py; python tools/trace_energy.py --streams 4; python tools/trace_energy.py --dtype bf16 --streams 4; } > gpurun_out/trace.log 2>&1
python - <<'PY'
import json
for f in ("bench_r01.json", "bench_r01_bf16.json", "bench_r01_reference.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], (d.get("roofline") or {}).get("frac"), d.get("clocks"), str(d.get("aux"))[:600])
    except Exception as e:
        print(f, "unreadable:", e)
PY
