#!/bin/bash
# Round 2, call AP (1 GPU): the driver sequence on the final code (smoke, parity suite, bench line, reference arm).
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_ap.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_ap.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["roofline"]["multi_stream_frac"], d["e2e"]["value"], d["e2e"]["all_region_ms_per_step"], d["e2e"]["frac_of_copy_ceiling"], d["clocks"])
a = d["aux"]
print("bf16", a["k1_bf16"]["single_stream_frac"], a["k1_bf16"]["multi_stream"]["frac"], "mixed", a["k1_bf16_x0f32"]["single_stream_frac"], a["k1_bf16_x0f32"]["multi_stream"]["frac"])
print("l2", a["k1_l2_resident"])
print("m32", a["k1_m32_bf16"]["speedup_1_stream"], a["k1_m32_bf16"]["speedup_4_streams"], a["k1_m32_bf16"]["tensor_core"]["us_per_launch_1_stream"])
print("dit", {k: round(a["dit_train"][k]["img_per_s"]) for k in ("bf16", "tf32", "fp32")}, "sampler", round(a["sampler"]["steps20"]["samples_per_s"]), "mmd", a["rbf_mmd2"]["ms"])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r02_reference.json
