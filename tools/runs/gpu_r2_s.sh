#!/bin/bash
# Round 2, call S (2 GPUs): data-parallel launcher forms: parity vs the global-batch recipe + DiT step time.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in split graph_nccl eager; do
  echo "== mode $mode"
  timeout 300 $TR --master-port $((29600 + RANDOM % 200)) tools/dp_overlap.py --mode $mode --steps 100 2> gpurun_out/dp_overlap_$mode.err | tail -1 | tee -a gpurun_out/dp_overlap.jsonl
  echo "rc=${PIPESTATUS[0]}"
  tail -3 gpurun_out/dp_overlap_$mode.err | cut -c1-300
done
