#!/usr/bin/env python
"""img/s of the DP DiT training step (BASELINE config 4) for a few launcher settings; run under torchrun for N > 1."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ddm_b200 import launcher

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--configs", default="--precision bf16;--precision bf16 --no-cuda-graph;--precision tf32")
a = ap.parse_args()
world, rank, dev = launcher.init_distributed()
for cfg in a.configs.split(";"):
    targs = launcher.build_parser().parse_args(["--synthetic"] + cfg.split())
    r = launcher.measure_throughput(targs, dev, world, steps=a.steps, warmup=3)
    if rank == 0:
        print(json.dumps({"cfg": cfg, **r}), flush=True)
    torch.cuda.empty_cache()
if world > 1:
    torch.distributed.destroy_process_group()
