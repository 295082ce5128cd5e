#!/bin/bash
# Round 2, call AU: full parity suite + smoke after the B-aware cluster choice (the 8-warp opt-in keeps whole rows).
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_au.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_au.log
