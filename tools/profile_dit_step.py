#!/usr/bin/env python
"""Kernel-level breakdown of one data-parallel DiT training step (BASELINE config 4) with torch.profiler.

    python tools/profile_dit_step.py [--precision bf16] [--steps 3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from ddm_b200 import launcher

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--extra", default="", help="extra launcher flags, space separated")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
targs = launcher.build_parser().parse_args(["--synthetic", "--precision", a.precision] + a.extra.split())
tr = launcher.Trainer(targs, dev, 1)
x0 = tr.synthetic_batch()
for _ in range(5):
    tr.step(x0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    tr.step(x0)
e1.record()
e1.synchronize()
print(f"step: {e0.elapsed_time(e1) / 10:.2f} ms  ({targs.batch * 10 / (e0.elapsed_time(e1) * 1e-3):.0f} img/s)")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        tr.step(x0)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
