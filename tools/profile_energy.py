#!/usr/bin/env python
"""Minimal driver for ncu: a handful of fused energy-score launches at the headline shape."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ddm_b200 import _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=128)
ap.add_argument("--m", type=int, default=8)
ap.add_argument("--D", type=int, default=3072)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--tune", default="")
a = ap.parse_args()
L = _cabi.lib()
for kv in filter(None, a.tune.split(",")):
    k, v = kv.split("=")
    _cabi.set_tuning(k, int(v))
dev = torch.device("cuda:0")
td = torch.float32 if a.dtype == "f32" else torch.bfloat16
fn = getattr(L, f"dddm_energy_fused_{a.dtype}")
sets = []
for s in range(a.iters):
    x0 = torch.randn(a.B, a.D, device=dev).clamp(-1, 1)
    xh = x0[:, None] + 0.05 * torch.randn(a.B, a.m, a.D, device=dev)
    sets.append((xh.to(td), x0.to(td), torch.empty(a.B, a.m, a.D, dtype=td, device=dev), torch.zeros(4, device=dev),
                 torch.full((1,), 0.5 * a.B, device=dev),
                 torch.zeros(L.dddm_energy_workspace_bytes(a.B, a.m), dtype=torch.uint8, device=dev)))
torch.cuda.synchronize()
cs = torch.cuda.current_stream().cuda_stream
for xh, x0, g, out, w, ws in sets:
    _cabi.check(fn(xh.data_ptr(), x0.data_ptr(), w.data_ptr(), 1.0 / a.B, g.data_ptr(), out.data_ptr(), ws.data_ptr(),
                   a.B, a.m, a.D, 0.1, 1.0, cs))
torch.cuda.synchronize()
print("ok", _cabi.describe_energy(a.B, a.m, a.D, a.dtype), sets[-1][3].tolist())
