#!/usr/bin/env python
"""Achieved HBM bandwidth of the elementwise kernels around the energy score at their BASELINE shapes (CUDA graph of
back-to-back launches over rotating buffers larger than L2 — inputs AND outputs distinct per launch — CUDA events).
One JSON line per kernel.  `bench.py` (`aux.elementwise`) runs the same measurement with 8x L2 of rotating sets."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import argparse

from ddm_b200 import _cabi, ops

ap = argparse.ArgumentParser()
ap.add_argument("--pdl", type=int, default=1, help="programmatic dependent launch of the streaming kernels (tuning energy.pdl)")
ap.add_argument("--only", default="", help="comma-separated prefixes of the kernels to time (K2,K2c,K3,K4,K6); default all")
a = ap.parse_args()
_cabi.set_tuning("energy.pdl", a.pdl)
ONLY = [p_ for p_ in a.only.split(",") if p_]

dev = torch.device("cuda:0")
PEAK = 6452.5


def bench(name, make, call, nbytes, nsets):
    if ONLY and name.split()[0] not in ONLY:
        return
    nsets = max(nsets, -(-8 * 126 * 2**20 // nbytes)) if nbytes > 2**20 else nsets  # 8 x L2 of rotating sets
    sets = [make(i) for i in range(nsets)]
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        for s in sets:
            call(s)
    stream.synchronize()
    g = torch.cuda.CUDAGraph()
    reps = 1
    keep = []  # the ops allocate their outputs: kept alive so that every launch writes its own, HBM-cold block
    with torch.cuda.graph(g, stream=stream):
        for s in sets:
            keep.append(call(s))
    ts = []
    with torch.cuda.stream(stream):
        g.replay()
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.replay()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / (reps * nsets))
    us = sorted(ts)[2]
    print(json.dumps({"pdl": a.pdl, "kernel": name, "us": round(us, 2), "algorithmic_MB": round(nbytes / 1e6, 2),
                      "GBps": round(nbytes / us / 1e3, 1), "frac_of_measured_hbm_peak": round(nbytes / us / 1e3 / PEAK, 3)}),
          flush=True)


B, m, C, H, W = 128, 8, 3, 32, 32
D = C * H * W
f32, bf = torch.float32, torch.bfloat16

# K2: forward marginal + m-fold expansion (config 4 shape)
bench("K2 forward_marginal_expand f32 [128 -> 1024 x 3072]",
      lambda i: (torch.rand(B, C, H, W, device=dev), torch.rand(B, device=dev), torch.randn(B, C, H, W, device=dev)),
      lambda s: ops.forward_marginal_expand(s[0], s[1], s[2], m, False), (2 * B * D + B * m * D) * 4 + 4 * B, 24)
# K2c: concat input in bf16 + x0 tokens
bench("K2c forward_marginal_concat f32 -> bf16 x6 [1024 x 6 x 32 x 32] + x0 tokens",
      lambda i: (torch.rand(B, C, H, W, device=dev), torch.rand(B, device=dev), torch.randn(B, C, H, W, device=dev),
                 torch.randn(B, m, C, H, W, device=dev)),
      lambda s: ops.forward_marginal_concat(s[0], s[1], s[2], s[3], True, 4),
      (2 * B * D + B * m * D) * 4 + 2 * B * m * D * 2 + B * D * 4, 16)
# K3: Algorithm-2 update, 1024 samples (config 5 on one GPU)
N = 1024
bench("K3 bridge_step f32 [1024 x 3072]",
      lambda i: (torch.randn(N, C, H, W, device=dev), torch.randn(N, C, H, W, device=dev), torch.randn(N, C, H, W, device=dev),
                 torch.tensor([0.45], device=dev), torch.tensor([0.5], device=dev)),
      lambda s: ops.bridge_step(s[0], s[1], s[2], s[3], s[4], 1.0), 4 * N * D * 4, 12)
# K4: logistic weight sum
bench("K4 sigmoid_weight_sum [128]", lambda i: (torch.rand(B, device=dev),), lambda s: ops.sigmoid_weight_sum(s[0], 0.0),
      B * 8 + 4, 8)
# K6: LayerNorm / column sums at the DiT shape
T_, Cn = 65536, 384
mk = lambda i: (torch.randn(T_, Cn, device=dev, dtype=bf), torch.ones(Cn, device=dev, dtype=bf), torch.zeros(Cn, device=dev, dtype=bf))
bench("K6 layer_norm fwd bf16 [65536 x 384]", mk, lambda s: ops.layer_norm(s[0], s[1], s[2], 1e-5), 2 * T_ * Cn * 2 + T_ * 8, 6)
mkb = lambda i: (torch.randn(T_, Cn, device=dev, dtype=bf), torch.randn(T_, Cn, device=dev, dtype=bf),
                 torch.zeros(T_, device=dev), torch.ones(T_, device=dev), torch.ones(Cn, device=dev, dtype=bf))
bench("K6 layer_norm bwd bf16 [65536 x 384]", mkb, lambda s: ops.layer_norm_bwd(s[0], s[1], s[2], s[3], s[4]),
      3 * T_ * Cn * 2 + T_ * 8, 6)
bench("K6 colsum bf16 [65536 x 384]", lambda i: (torch.randn(T_, Cn, device=dev, dtype=bf),), lambda s: ops.colsum(s[0]),
      T_ * Cn * 2, 8)
bench("K6 colsum bf16 [65536 x 1536]", lambda i: (torch.randn(T_, 1536, device=dev, dtype=bf),), lambda s: ops.colsum(s[0]),
      T_ * 1536 * 2, 4)
