#!/bin/bash
# Round 2, call AG (1 GPU): final validation on the final code — smoke, GPU suite, the driver's bench command and reference
# arm, then the ncu captures (each after the same command ran plainly): K1 fp32 full set, launch list of the bench command.
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_ag.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_ag.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["roofline"]["multi_stream_frac"], d["e2e"]["value"], d["e2e"]["all_region_ms_per_step"], d["e2e"]["frac_of_copy_ceiling"], d["clocks"])
a = d["aux"]
print("bf16", a["k1_bf16"]["single_stream_frac"], a["k1_bf16"]["multi_stream"]["frac"], "mixed", a["k1_bf16_x0f32"]["single_stream_frac"], a["k1_bf16_x0f32"]["multi_stream"]["frac"])
print("m32", a["k1_m32_bf16"]["speedup_1_stream"], a["k1_m32_bf16"]["speedup_4_streams"], a["k1_m32_bf16"]["tensor_core"]["us_per_launch_1_stream"])
print("copy", a["single_launch_copy_ceiling"])
for k, v in a["elementwise"].items():
    print(k, round(v["us_per_launch"], 2), round(v["roofline"]["frac"], 3))
print("dit", {k: round(a["dit_train"][k]["img_per_s"]) for k in ("bf16", "tf32", "fp32")}, "sampler", round(a["sampler"]["steps20"]["samples_per_s"]), "mmd", a["rbf_mmd2"]["ms"])
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_reference.json 2>/dev/null ) 2>&1 | grep real; cut -c1-300 gpurun_out/bench_r02_reference.json
P="python tools/profile_energy.py"
$P > gpurun_out/plain_k1_f32.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_fused_smem -s 4 -c 3 -f -o gpurun_out/prof_r02_k1_f32 $P > gpurun_out/ncu_k1_f32.log 2>&1
$P --dtype bf16 > gpurun_out/plain_k1_bf16.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:energy_fused_smem -s 4 -c 3 -f -o gpurun_out/prof_r02_k1_bf16 $P --dtype bf16 > gpurun_out/ncu_k1_bf16.log 2>&1
Bn="python bench.py --steps 200 --warmup 3 --cpu-seconds 0 --dit-steps 0 --sampler-samples 0 --mmd-samples 0 --no-elementwise --e2e-steps 3"
$Bn > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r02.csv $Bn > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_*.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench_r02.csv
