"""Timeline of the host-buffer pipeline (diagnostics for bench.py's `e2e`): the same H2D -> K4 + K1 -> D2H pipeline
as csrc/host_session.cu, rebuilt with torch streams and TIMING events so every copy and kernel of every step can be
placed on one time axis, plus the C-ABI session itself timed by the wall clock on the same box.

    python tools/e2e_trace.py [--steps 20] [--slots 4] [--host-sets 4] [--h2d-chunks 1] [--d2h-chunks 1]
"""
from __future__ import annotations

import argparse
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--host-sets", type=int, default=4)
    ap.add_argument("--h2d-chunks", type=int, default=1)
    ap.add_argument("--d2h-chunks", type=int, default=1)
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args()
    import torch

    from ddm_b200 import _cabi

    L = _cabi.lib()
    dev = torch.device("cuda:0")
    B, M, D = 128, 8, 3072
    nx, n0 = B * M * D * 4, B * D * 4
    off_t = nx + n0
    in_bytes, out_bytes = off_t + 4 * B, nx + 256
    gen = torch.Generator().manual_seed(0)
    host = []
    for s in range(a.host_sets):
        pin = torch.empty(in_bytes, dtype=torch.uint8).pin_memory()
        pin[:off_t] = (0.1 * torch.randn(off_t // 4, generator=gen)).view(torch.uint8)
        pin[off_t:] = torch.rand(B, generator=gen).view(torch.uint8)
        host.append((pin, torch.empty(out_bytes, dtype=torch.uint8).pin_memory()))
    slots = []
    for s in range(a.slots):
        slots.append({"in": torch.empty(in_bytes, dtype=torch.uint8, device=dev),
                      "out": torch.empty(out_bytes, dtype=torch.uint8, device=dev),
                      "wsum": torch.empty(1, device=dev),
                      "ws": torch.zeros(L.dddm_energy_workspace_bytes(B, M), dtype=torch.uint8, device=dev),
                      "done": None})
    s_in, s_run, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def chunks(n, k):
        step = (n + k - 1) // k
        step = (step + 255) // 256 * 256
        return [(o, min(step, n - o)) for o in range(0, n, step)]

    def run(n, record):
        rows = []
        for i in range(n):
            k = slots[i % a.slots]
            pin, pout = host[i % a.host_sets]
            if k["done"] is not None:
                k["done"].synchronize()
            e = [ev() for _ in range(6)]
            with torch.cuda.stream(s_in):
                e[0].record()
                for o, c in chunks(in_bytes, a.h2d_chunks):
                    k["in"][o:o + c].copy_(pin[o:o + c], non_blocking=True)
                e[1].record()
            s_run.wait_event(e[1])
            with torch.cuda.stream(s_run):
                e[2].record()
                p = k["in"].data_ptr()
                _cabi.check(L.dddm_sigmoid_weight_sum_f32(p + off_t, 0.0, None, k["wsum"].data_ptr(), B, s_run.cuda_stream))
                _cabi.check(L.dddm_energy_fused_f32(p, p + nx, k["wsum"].data_ptr(), 1.0 / B, k["out"].data_ptr(),
                                                    k["out"].data_ptr() + nx, k["ws"].data_ptr(), B, M, D, 0.1, 1.0,
                                                    s_run.cuda_stream))
                e[3].record()
            s_out.wait_event(e[3])
            with torch.cuda.stream(s_out):
                e[4].record()
                for o, c in chunks(nx + 16, a.d2h_chunks):
                    pout[o:o + c].copy_(k["out"][o:o + c], non_blocking=True)
                e[5].record()
            k["done"] = e[5]
            rows.append(e)
        torch.cuda.synchronize(dev)
        return rows

    run(8, False)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    rows = run(a.steps, True)
    wall = time.perf_counter() - t0
    base = rows[0][0]
    print(f"torch-stream pipeline: slots={a.slots} host_sets={a.host_sets} h2d_chunks={a.h2d_chunks} d2h_chunks={a.d2h_chunks}: "
          f"{1e3 * wall / a.steps:.4f} ms/step wall ({B * a.steps / wall:.0f} rows/s)")
    if not a.quiet:
        print("step  h2d_start  h2d_end  (GB/s)   k_start  k_end   d2h_start  d2h_end  (GB/s)   [ms from the first upload]")
        for i, e in enumerate(rows):
            t = [base.elapsed_time(x) for x in e]
            print(f"{i:3d}  {t[0]:8.3f} {t[1]:8.3f}  {in_bytes / (t[1] - t[0]) / 1e6:6.1f}  {t[2]:8.3f} {t[3]:8.3f}  "
                  f"{t[4]:8.3f} {t[5]:8.3f}  {(nx + 16) / (t[5] - t[4]) / 1e6:6.1f}")

    # the C-ABI session on the same box, same bytes, wall clock
    sess = L.dddm_session_create(B, M, D, 0, 0)
    sz = [ctypes.c_size_t() for _ in range(5)]
    _cabi.check(L.dddm_session_packed_layout(sess, *[ctypes.addressof(v) for v in sz]))
    s_in_bytes, x0_off, t_off, s_out_bytes, out_off = [int(v.value) for v in sz]
    assert s_in_bytes <= in_bytes + 256 and x0_off == nx and t_off == off_t and out_off == nx, (s_in_bytes, x0_off, t_off, out_off)

    def sess_run(n):
        for i in range(n):
            pin, pout = host[i % a.host_sets]
            _cabi.check(L.dddm_session_enqueue_host(sess, pin.data_ptr(), pin.data_ptr() + x0_off, pin.data_ptr() + t_off,
                                                    0.0, 0.1, 1.0, pout.data_ptr(), pout.data_ptr() + out_off))
        _cabi.check(L.dddm_session_wait(sess))

    sess_run(8)
    for n in (a.steps, a.steps, 10 * a.steps):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        sess_run(n)
        dt = time.perf_counter() - t0
        print(f"C-ABI session, {n} steps: {1e3 * dt / n:.4f} ms/step ({B * n / dt:.0f} rows/s)")
    # the same with WRITE-COMBINED pinned input buffers (dddm_host_alloc_input): the CPU only writes them
    wc = []
    for s in range(a.host_sets):
        ptr = L.dddm_host_alloc_input(in_bytes)
        assert ptr, "dddm_host_alloc_input failed"
        ctypes.memmove(ptr, host[s][0].data_ptr(), in_bytes)
        wc.append(ptr)

    def sess_run_wc(n):
        for i in range(n):
            pin, pout = wc[i % a.host_sets], host[i % a.host_sets][1]
            _cabi.check(L.dddm_session_enqueue_host(sess, pin, pin + x0_off, pin + t_off, 0.0, 0.1, 1.0, pout.data_ptr(),
                                                    pout.data_ptr() + out_off))
        _cabi.check(L.dddm_session_wait(sess))

    ref_out = host[0][1][out_off:out_off + 16].clone()
    sess_run_wc(8)
    assert torch.equal(host[0][1][out_off:out_off + 16], ref_out), "write-combined inputs changed the result"
    for n in (a.steps, a.steps, 10 * a.steps):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        sess_run_wc(n)
        dt = time.perf_counter() - t0
        print(f"C-ABI session, write-combined inputs, {n} steps: {1e3 * dt / n:.4f} ms/step ({B * n / dt:.0f} rows/s)")
    for ptr in wc:
        L.dddm_host_free(ptr)
    L.dddm_session_destroy(sess)


if __name__ == "__main__":
    main()
