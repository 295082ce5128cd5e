#!/bin/bash
# Round 2, call AD: polled finish (default) + bulk-store pass 2 (A/B), short bench line (e2e on this box), e2e tool.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_energy.py -m gpu -q -x > gpurun_out/pytest_gpu_ad.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_ad.log
echo "== f32 one stream: stores from registers (bulkst=1) vs bulk stores (bulkst=2)"
timeout 120 python tools/sweep_energy.py --streams 1 --configs "variant=3,bulkst=1;variant=3,bulkst=2;variant=3,bulkst=1;variant=3,bulkst=2"
echo "== f32 six streams"
timeout 120 python tools/sweep_energy.py --streams 6 --configs "variant=3,bulkst=1;variant=3,bulkst=2;variant=3,finish=1"
echo "== bf16 one / six streams"
timeout 120 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3,bulkst=1;variant=3,bulkst=2"
timeout 120 python tools/sweep_energy.py --streams 6 --dtype bf16 --configs "variant=3,bulkst=1;variant=3,bulkst=2"
echo "== trace (bulk stores)"
timeout 200 python tools/trace_energy.py --tune energy.bulkst=2 > gpurun_out/k1_trace_bulkst.log 2>&1; grep "us/launch\|->\|period\|row_finished\|pass2_done\|span" gpurun_out/k1_trace_bulkst.log | head -14
echo "== bench (no aux)"
timeout 600 python bench.py --steps 20 --warmup 5 --dit-steps 0 --sampler-samples 0 --mmd-samples 0 --no-elementwise > gpurun_out/bench_ad.json 2> gpurun_out/bench_ad.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_ad.json").read().strip().splitlines()[-1])
print("single", d["ms_per_step"], "frac", d["roofline"]["frac"], "multi", d["config"]["multi_stream"]["ms_per_step"])
print("e2e", json.dumps(d["e2e"])[:900])
PY
timeout 200 python tools/e2e_trace.py --quiet
