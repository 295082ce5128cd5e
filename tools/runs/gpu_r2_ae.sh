#!/bin/bash
# Round 2, call AE: store path per byte or per instruction (ubench); double-width stores by lane pairs in pass 2 (A/B).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_energy.py -m gpu -q -x > gpurun_out/pytest_gpu_ae.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_ae.log
timeout 120 tools/ubench/store_width > gpurun_out/store_width.log 2>&1; echo "store_width rc=$?"; tail -28 gpurun_out/store_width.log
echo "== f32 one stream: bulkst=1 (one store per step) vs 3 (wide stores)"
timeout 120 python tools/sweep_energy.py --streams 1 --configs "variant=3,bulkst=1;variant=3,bulkst=3;variant=3,bulkst=1;variant=3,bulkst=3"
echo "== f32 six streams"
timeout 120 python tools/sweep_energy.py --streams 6 --configs "variant=3,bulkst=1;variant=3,bulkst=3"
echo "== bf16 one / six streams"
timeout 120 python tools/sweep_energy.py --streams 1 --dtype bf16 --configs "variant=3,bulkst=1;variant=3,bulkst=3"
timeout 120 python tools/sweep_energy.py --streams 6 --dtype bf16 --configs "variant=3,bulkst=1;variant=3,bulkst=3"
echo "== trace (wide stores)"
timeout 200 python tools/trace_energy.py --tune energy.bulkst=3 > gpurun_out/k1_trace_wide.log 2>&1; grep "us/launch\|->\|period\|row_finished\|pass2_done\|span" gpurun_out/k1_trace_wide.log | head -14
timeout 200 python tools/trace_energy.py --dtype bf16 --tune energy.bulkst=3 > gpurun_out/k1_trace_wide_bf16.log 2>&1; grep "us/launch\|->\|period\|row_finished\|pass2_done\|span" gpurun_out/k1_trace_wide_bf16.log | head -14
