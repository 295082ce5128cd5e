#!/usr/bin/env python
"""In-kernel timeline of the fused energy-score kernel (dddm_set_trace_buffer): where do the
microseconds of one launch go?  Prints, per stage, the median / max over CTAs of the time since the
launch's first CTA entered, averaged over the launches of a CUDA graph, plus the gap between the end of
launch k and the start of launch k+1 on the same stream.

    python tools/trace_energy.py [--B 128 --m 8 --D 3072 --dtype f32 --streams 1 --tune energy.pdl=1]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ddm_b200 import _cabi

STAGES = ["entry", "inputs_ready(griddep)", "first_chunk_landed", "pass1_done", "coef_ready", "pass2_done",
          "tma_issued(ctrl)", "row_finished(ctrl)", "warp_reduce_done", "after_sync1", "coef_written", "last_warp_pass1_done"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=128)
    ap.add_argument("--m", type=int, default=8)
    ap.add_argument("--D", type=int, default=3072)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--launches", type=int, default=40)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--tune", default="")
    a = ap.parse_args()
    L = _cabi.lib()
    for kv in filter(None, a.tune.split(",")):
        k, v = kv.split("=")
        _cabi.set_tuning(k, int(v))
    dev = torch.device("cuda:0")
    td = torch.float32 if a.dtype == "f32" else torch.bfloat16
    fn = getattr(L, f"dddm_energy_fused_{a.dtype}")
    desc = _cabi.describe_energy(a.B, a.m, a.D, a.dtype)
    cluster = int(desc.split(" cluster=")[1].split()[0]) if " cluster=" in desc else 1
    sets = []
    for s in range(a.launches):
        g = torch.Generator().manual_seed(s)
        x0 = torch.randn(a.B, a.D, generator=g).clamp(-1, 1)
        xh = x0[:, None] + 0.05 * torch.randn(a.B, a.m, a.D, generator=g)
        sets.append((xh.to(td).to(dev), x0.to(td).to(dev), torch.empty(a.B, a.m, a.D, dtype=td, device=dev),
                     torch.zeros(4, device=dev), torch.full((1,), 0.5 * a.B, device=dev),
                     torch.zeros(L.dddm_energy_workspace_bytes(a.B, a.m), dtype=torch.uint8, device=dev),
                     torch.zeros(a.B * cluster * 16, dtype=torch.int64, device=dev)))
    stream = torch.cuda.Stream(dev)
    sides = [torch.cuda.Stream(dev) for _ in range(a.streams - 1)]
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=stream):
        main_s = torch.cuda.current_stream()
        for sd in sides:
            sd.wait_stream(main_s)
        for i, (xh, x0, gr, out, w, ws, tr) in enumerate(sets):
            st = main_s if i % a.streams == 0 else sides[i % a.streams - 1]
            _cabi.check(L.dddm_set_trace_buffer(tr.data_ptr()))
            _cabi.check(fn(xh.data_ptr(), x0.data_ptr(), w.data_ptr(), 1.0 / a.B, gr.data_ptr(), out.data_ptr(),
                           ws.data_ptr(), a.B, a.m, a.D, 0.1, 1.0, st.cuda_stream))
        for sd in sides:
            main_s.wait_stream(sd)
    _cabi.check(L.dddm_set_trace_buffer(None))
    with torch.cuda.stream(stream):
        for _ in range(3):
            graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        graph.replay()
        e1.record(stream)
        e1.synchronize()
    print(f"kernel: {desc}; streams={a.streams}; tune={a.tune or 'auto'}; "
          f"{e0.elapsed_time(e1) * 1e3 / a.launches:.2f} us/launch (events, with tracing on)")
    full = np.stack([s[6].cpu().numpy().reshape(a.B * cluster, 16) for s in sets]).astype(np.int64)
    tr = full[:, :, :12]  # [launch, cta, stage]
    if full.shape[0] > 2 and (full[2:, :, 13] > 0).all():  # SM cycle counter stamped next to inputs_ready / pass2_done: effective SM clock
        mhz = (full[2:, :, 13] - full[2:, :, 12]) / np.maximum(1, tr[2:, :, 5] - tr[2:, :, 1]) * 1e3
        print(f"SM clock between inputs_ready and pass2_done: median {np.median(mhz):.0f} MHz (p5 {np.percentile(mhz, 5):.0f}, "
              f"p95 {np.percentile(mhz, 95):.0f})")
    t0 = tr[:, :, 0].min(axis=1)  # first CTA entry of each launch
    rel = tr - t0[:, None, None]
    res = np.diff(np.unique(tr[:, :, 0].ravel()))
    print(f"globaltimer resolution (smallest step seen): {res[res > 0].min() if (res > 0).any() else 'n/a'} ns")
    print(f"{'stage':<26}{'median over CTAs (ns)':>24}{'max over CTAs (ns)':>22}   (mean over launches 2..)")
    for k, name in enumerate(STAGES):
        med = np.median(rel[2:, :, k], axis=1).mean()
        mx = rel[2:, :, k].max(axis=1).mean()
        print(f"{name:<26}{med:>24.0f}{mx:>22.0f}")
    end = tr[:, :, [5, 7]].max(axis=(1, 2))
    if a.streams == 1:
        gap = t0[1:] - end[:-1]
        print(f"gap: end of launch k -> first CTA entry of launch k+1: mean {gap[1:].mean():.0f} ns, min {gap[1:].min()}, "
              f"max {gap[1:].max()}")
        per = np.diff(t0)[1:]
        print(f"start-to-start period: mean {per.mean():.0f} ns")
    dur = end - t0
    print(f"launch span (first entry -> last stamp): mean {dur[2:].mean():.0f} ns")
    # per-CTA stage durations
    seg = [("entry->inputs_ready", 0, 1), ("inputs_ready->first_chunk", 1, 2), ("first_chunk->pass1_done", 2, 3),
           ("pass1_done->coef_ready", 3, 4), ("  pass1_done->warp_reduce_done", 3, 8), ("  warp_reduce->after_sync1", 8, 9),
           ("  after_sync1->coef_written", 9, 10), ("  coef_written->coef_ready(sync2)", 10, 4),
           ("  warp0 vs last warp pass1 exit", 3, 11), ("coef_ready->pass2_done", 4, 5)]
    if desc.startswith("tc<"):  # tensor-core kernel: its own stamp layout (energy_tc.cu TC_TRACE)
        names = ["entry", "inputs_ready", "tma_pass1_issued", "tma_pass2_issued", "gram_committed", "pass2_mmas_committed",
                 "conf_pass_done", "gram_in_smem", "coef_ready", "epilogue_done", "row_done", "row_published"]
        print("tensor-core kernel stamps, ns after inputs_ready (median / max over CTAs, mean over launches 2..):")
        for k, name in enumerate(names):
            dlt = tr[2:, :, k] - tr[2:, :, 1]
            print(f"  {name:<24}{np.median(dlt, axis=1).mean():>10.0f}{dlt.max(axis=1).mean():>10.0f}")
        return
    if desc.startswith("pipe<"):  # row-pipelined cluster kernel: slots 2-4 are row 0, slots 8-10 the cluster's last row
        seg = [("entry->inputs_ready", 0, 1), ("inputs_ready->row0_landed", 1, 2), ("row0_landed->row0_published", 2, 3),
               ("row0_published->row0_coef", 3, 4), ("inputs_ready->last_row_landed", 1, 8),
               ("last_row_landed->published", 8, 9), ("last_row_published->coef", 9, 10), ("last_row_coef->pass2_done", 10, 5),
               ("inputs_ready->pass2_done", 1, 5)]
    for name, i, j in seg:
        d = (tr[2:, :, j] - tr[2:, :, i])
        print(f"  {name:<28} median {np.median(d):>7.0f} ns   p95 {np.percentile(d, 95):>7.0f} ns")


if __name__ == "__main__":
    main()
