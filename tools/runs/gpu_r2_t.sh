#!/bin/bash
# Round 2, call T (1 GPU): the driver's bench invocation, then ncu captures (each after the same command ran plainly).
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r02.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02.json").read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "roofline", "e2e", "clocks", "gpu_launches")}, indent=1)[:2500])
    a = d.get("aux", {})
    for k in ("k1_bf16", "k1_bf16_x0f32", "k1_m32_bf16", "single_launch_copy_ceiling", "elementwise", "sampler", "rbf_mmd2"):
        print(k, json.dumps(a.get(k))[:1500])
    print("dit", json.dumps({k: v for k, v in a.get("dit_train", {}).items() if k in ("bf16", "tf32", "fp32")})[:900])
except Exception as e:
    print("parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_r02_reference.json
P="python tools/profile_energy.py"
$P > gpurun_out/plain_k1_f32.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_fused_smem -s 4 -c 3 -f -o gpurun_out/prof_r02_k1_f32 $P > gpurun_out/ncu_k1_f32.log 2>&1
$P --dtype bf16 --m 32 --iters 6 > gpurun_out/plain_tc_m32.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_tc_kernel -s 2 -c 2 -f -o gpurun_out/prof_r02_tc_m32 $P --dtype bf16 --m 32 --iters 6 > gpurun_out/ncu_tc_m32.log 2>&1
Bn="python bench.py --steps 200 --warmup 3 --cpu-seconds 0 --dit-steps 0 --sampler-samples 0 --mmd-samples 0 --no-elementwise --e2e-steps 3"
$Bn > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r02.csv $Bn > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_*.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench_r02.csv
