#!/bin/bash
# Round 2, call X: final checks: smoke(), GPU suite, tensor-core kernel timeline after the epilogue reorder.
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_x.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_x.log
timeout 300 python tools/trace_energy.py --dtype bf16 --m 32 --tune "energy.variant=7" | grep -A12 "tensor-core kernel stamps\|us/launch" | grep "us/launch\|conf_pass_done\|gram_in_smem\|coef_ready\|epilogue_done\|row_done"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --m 32 --configs "variant=7"
timeout 300 python tools/sweep_energy.py --streams 4 --dtype bf16 --m 32 --configs "variant=7"
timeout 300 python tools/sweep_energy.py --streams 1 --dtype bf16 --m 16 --configs "variant=7"
