#!/bin/bash
# Round 2, call AC: polled row-slot finish of the TMA-staged kernel (A/B against the ticket), e2e pipeline timeline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_energy.py -m gpu -q -x > gpurun_out/pytest_gpu_ac.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_ac.log
echo "== f32 one stream: ticket (finish=1) vs polled (finish=2)"
timeout 300 python tools/sweep_energy.py --streams 1 --configs "variant=3,finish=1;variant=3,finish=2;variant=3,finish=1;variant=3,finish=2"
echo "== f32 six streams"
timeout 300 python tools/sweep_energy.py --streams 6 --configs "variant=3,finish=1;variant=3,finish=2"
echo "== bf16 one / six streams"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3,finish=1;variant=3,finish=2"
timeout 300 python tools/sweep_energy.py --streams 6 --dtype bf16 --configs "variant=3,finish=1;variant=3,finish=2"
echo "== B = 1024 one stream"
timeout 300 python tools/sweep_energy.py --streams 1 --B 1024 --configs "variant=3,finish=1;variant=3,finish=2"
echo "== trace (polled)"
timeout 200 python tools/trace_energy.py > gpurun_out/k1_trace_polled.log 2>&1; grep "us/launch\|->\|period\|row_finished\|pass2_done\|span" gpurun_out/k1_trace_polled.log | head -24
echo "== e2e timeline"
timeout 200 python tools/e2e_trace.py > gpurun_out/e2e_trace_a.log 2>&1; tail -30 gpurun_out/e2e_trace_a.log
timeout 200 python tools/e2e_trace.py --quiet --host-sets 1
timeout 200 python tools/e2e_trace.py --quiet --h2d-chunks 4 --d2h-chunks 4
timeout 200 python tools/e2e_trace.py --quiet --slots 8 --host-sets 8
