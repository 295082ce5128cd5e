#!/usr/bin/env python
"""us per launch of the split pair K1b (dddm_energy_terms_fwd / _bwd) at BASELINE config 2, next to the fused K1."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ddm_b200 import _cabi

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="f32")
ap.add_argument("--streams", type=int, default=4)
ap.add_argument("--m", type=int, default=8)
a = ap.parse_args()
L = _cabi.lib()
ap2 = None
B, m, D = 128, a.m, 3072
dev = torch.device("cuda:0")
td = torch.float32 if a.dtype == "f32" else torch.bfloat16
esz = 4 if a.dtype == "f32" else 2
P = m + m * (m - 1) // 2
nsets = 24
sets = []
for s in range(nsets):
    g = torch.Generator().manual_seed(s)
    x0 = torch.randn(B, D, generator=g).clamp(-1, 1)
    xh = x0[:, None] + 0.05 * torch.randn(B, m, D, generator=g)
    sets.append(dict(xh=xh.to(td).to(dev), x0=x0.to(td).to(dev), gx=torch.empty(B, m, D, dtype=td, device=dev),
                     dist=torch.empty(B, P, device=dev), out=torch.zeros(2, device=dev), gc=torch.ones(1, device=dev),
                     gi=torch.full((1,), -0.07, device=dev),
                     ws=torch.zeros(L.dddm_energy_workspace_bytes(B, m), dtype=torch.uint8, device=dev)))
fwd = getattr(L, f"dddm_energy_terms_fwd_{a.dtype}")
bwd = getattr(L, f"dddm_energy_terms_bwd_{a.dtype}")


def l_fwd(s, cs):
    _cabi.check(fwd(s["xh"].data_ptr(), s["x0"].data_ptr(), s["dist"].data_ptr(), s["out"].data_ptr(), s["ws"].data_ptr(),
                    B, m, D, 0.1, cs))


def l_bwd(s, cs):
    _cabi.check(bwd(s["xh"].data_ptr(), s["x0"].data_ptr(), s["dist"].data_ptr(), s["gc"].data_ptr(), s["gi"].data_ptr(),
                    s["gx"].data_ptr(), None, B, m, D, 0.1, cs))


def timeit(fn, nstreams):
    stream = torch.cuda.Stream(dev)
    sides = [torch.cuda.Stream(dev) for _ in range(nstreams - 1)]
    graph = torch.cuda.CUDAGraph()
    n = nsets * 10
    with torch.cuda.graph(graph, stream=stream):
        main = torch.cuda.current_stream()
        for sd in sides:
            sd.wait_stream(main)
        for i in range(n):
            st = main if i % nstreams == 0 else sides[i % nstreams - 1]
            fn(sets[i % nsets], st.cuda_stream)
        for sd in sides:
            main.wait_stream(sd)
    with torch.cuda.stream(stream):
        for _ in range(3):
            graph.replay()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(4):
                graph.replay()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / (4 * n))
    return sorted(ts)[2]


for s in sets:
    l_fwd(s, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
bf, bb = (B * m * D + B * D) * esz, (2 * B * m * D + B * D) * esz
for ns in (1, a.streams):
    tf, tb = timeit(l_fwd, ns), timeit(l_bwd, ns)
    print(f"{a.dtype} streams={ns}: fwd {tf:.2f} us ({bf / tf / 1e3:.0f} GB/s)  bwd {tb:.2f} us ({bb / tb / 1e3:.0f} GB/s)  "
          f"pair {tf + tb:.2f} us", flush=True)
