#!/bin/bash
# Round 2, call AH: host-buffer session with write-combined pinned input buffers (dddm_host_alloc_input) against plain pinned ones.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_step.py -m gpu -q -x -k "session or host" 2>&1 | tail -2
timeout 200 python tools/e2e_trace.py --quiet
timeout 200 python tools/e2e_trace.py --quiet
