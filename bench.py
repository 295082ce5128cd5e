#!/usr/bin/env python
"""bench.py — headline benchmark of the DDDM hot path on B200 (contract: see repo README / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f32|bf16]

One "step" = one pass of the hot path over one minibatch: the fused energy-score forward+backward
(K1) on synthetic CIFAR-shaped draws B=128, m=8, D=3072 (BASELINE.json configs[1]).  Inputs are
resident in HBM for `value`; `e2e` goes through the C-ABI host-buffer session (H2D + kernels + D2H).
Under torchrun (N > 1) every rank processes its own shard of the global batch (rows are
independent; no data-path collective; weak scaling) and rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B, M, D = 128, 8, 3072
BETA, LAM, W_BIAS = 0.1, 1.0, 0.0
METRIC = "energy-score fwd+bwd rows/s"
L2_BYTES = 126 * 1024 * 1024


def workload_name(dtype: str) -> str:
    return f"isolated energy-score loss fwd+bwd, synthetic CIFAR-shaped draws B={B} m={M} D={D} {dtype} beta={BETA}"


def make_inputs(seed: int, dtype, m: int = 0):
    """SURVEY.md §8(d) 'late' regime (the precision-critical one): x0 in CIFAR range, draws x0 + 0.05 noise."""
    import torch

    m = m or M
    gen = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, D, generator=gen).clamp(-1, 1)
    xh = x0[:, None, :] + 0.05 * torch.randn(B, m, D, generator=gen)
    t = torch.rand(B, generator=gen)
    return xh.to(dtype), x0.to(dtype), t


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs: NVML every 5 ms when `pynvml` is importable
    (the region is ~100 ms), else `nvidia-smi` every 100 ms."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
    NVML_BITS = (0x8, 0x40, 0x20, 0x4)  # nvmlClocksEventReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.source = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml, self.source = pynvml, "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM))
        try:
            bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
        except Exception:
            bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
        self.rows.append([str(sm), str(self._max), "0"] + ["Active" if bits & b else "Not Active" for b in self.NVML_BITS])

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    parts = [p.strip() for p in out.strip().split(",")]
                    if len(parts) >= 7:
                        self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.005 if self._nvml is not None else 0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = [n for k, n in enumerate(self.NAMES) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows), "source": self.source}


def cpu_port_rate(seconds: float, dtype_name: str):
    """The reference's CPU path (eager PyTorch + autograd; oracle/torch_port.py) on the host cores:
    fwd+bwd of the same workload, repeated for about `seconds`; returns (rows/s, iterations, threads)."""
    import torch

    from oracle import torch_port

    xh, x0, t = make_inputs(0, torch.float32)
    w = torch_port.sigmoid_weight(t, W_BIAS).mean()
    torch_port.energy_fwd_bwd(xh, x0, w, BETA, LAM)  # warm-up
    n, t0 = 0, time.perf_counter()
    best = float("inf")
    while True:
        a = time.perf_counter()
        torch_port.energy_fwd_bwd(xh, x0, w, BETA, LAM)
        best = min(best, time.perf_counter() - a)
        n += 1
        if time.perf_counter() - t0 >= seconds and n >= 3:
            break
    mean = (time.perf_counter() - t0) / n
    return B / mean, n, torch.get_num_threads(), mean, best


def bind_to_gpu_numa_node(index: int) -> str:
    """Pin this process to the CPUs NVML reports as local to the GPU, BEFORE any pinned host buffer is allocated:
    with 8 ranks on a two-socket host, pinned pages that land on the far socket halve the e2e copy rate."""
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64 if (os.cpu_count() or 0) > 0 else 1)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            return f"bound to {len(target)} of {len(allowed)} allowed CPUs local to GPU {phys}"
        return f"no narrower local CPU set ({len(cpus)} local, {len(allowed)} allowed)"
    except Exception as exc:  # noqa: BLE001 - affinity is an optimisation, never a requirement
        return f"not bound ({type(exc).__name__})"


def gpu_eager_rate(dev, iters: int = 30):
    """A second, tougher baseline: the reference's own eager PyTorch path (the port in
    oracle/torch_port.py, same ops as dddm/losses.py + autograd) run on the SAME B200.  Reported, never shipped."""
    import torch

    from oracle import torch_port

    xh, x0, t = make_inputs(0, torch.float32)
    xh, x0, t = xh.to(dev), x0.to(dev), t.to(dev)
    w = torch_port.sigmoid_weight(t, W_BIAS).mean()
    for _ in range(3):
        torch_port.energy_fwd_bwd(xh, x0, w, BETA, LAM)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        torch_port.energy_fwd_bwd(xh, x0, w, BETA, LAM)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"value": B / (ms * 1e-3), "unit": "rows/s", "ms_per_step": ms, "kind": "port-on-gpu",
            "sample": f"{iters} fwd+bwd passes of oracle/torch_port.py (eager PyTorch ops + autograd, as dddm/losses.py) on "
                      f"the same B200, fp32, inputs resident"}


def run_reference(args) -> None:
    """--impl reference: the reference's own CPU implementation of the path (oracle port: the
    reference is pure Python/PyTorch and /root/reference is absent on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import torch_port

    # torchrun exports OMP_NUM_THREADS=1 for N > 1: the reference arm uses every host core it is allowed to
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    xh, x0, t = make_inputs(0, torch.float32)
    w = torch_port.sigmoid_weight(t, W_BIAS).mean()
    # bounded: each step is ~0.1-0.3 s of CPU work; cap the whole run at a few minutes
    steps = max(1, min(args.steps, 200))
    warmup = max(3, min(args.warmup, 20))  # the driver's W is honoured (each CPU step is ~40 ms: 20 of them stay bounded)
    for _ in range(warmup):
        torch_port.energy_fwd_bwd(xh, x0, w, BETA, LAM)
    t0 = time.perf_counter()
    for _ in range(steps):
        torch_port.energy_fwd_bwd(xh, x0, w, BETA, LAM)
    dt = time.perf_counter() - t0
    value = B * steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name("f32"), "rows_per_gpu": B, "global_rows": B,
                   "parallelism": "host CPU (rank 0 only)", "l2_policy": "n/a (CPU)",
                   "launch": f"eager PyTorch + autograd on {torch.get_num_threads()} host threads",
                   # the same key set as the GPU arm's config (the driver compares the two lines key by key)
                   "single_stream_ms_per_step": 1e3 * dt / steps, "single_stream_rows_per_s": value,
                   "multi_stream": {"streams": 1, "ms_per_step": 1e3 * dt / steps, "rows_per_s": value,
                                    "note": "n/a on the CPU: one pass at a time over all host threads"},
                   "timed_region": f"exactly {steps} steps, wall clock", "kernel": "oracle/torch_port.py", "tuning": "n/a"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{steps} full fwd+bwd passes over the B={B} batch (oracle/torch_port.py, eager "
                                   f"PyTorch + autograd as in the reference), os.cpu_count()={os.cpu_count()}"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


AUX_PARTIAL: dict = {}  # what run_aux has finished so far (reported if its watchdog fires)


def run_aux(args, dev, world, rank, L, peak) -> dict:
    """Auxiliary, outside the timed region: the other dtype of BASELINE config 2, the elementwise kernels' rooflines,
    BASELINE configs 4 (DP DiT training img/s, + multi-rank parity on hardware) and 5 (Algorithm-2 sampler), rbf_mmd2."""
    import torch
    import torch.distributed as dist

    from ddm_b200 import _cabi

    aux = AUX_PARTIAL

    # -- BASELINE config 2, second half: the same isolated loss in bf16 (all-bf16 and the mixed bf16-draws/fp32-data entry)
    def k1_block(dtype_name, x0_f32=False):
        kb = K1Bench(L, dev, dtype_name, rank, world, args.streams, args.no_graph, x0_f32=x0_f32)
        K = 480
        with torch.cuda.stream(kb.stream):
            kb.run_serial(K)
            kb.run_multi(K)
            kb.stream.synchronize()
            ser, _ = kb.timed(kb.run_serial, K, 5)
            mul, _ = kb.timed(kb.run_multi, K, 3)
        if world > 1:
            tt = torch.tensor([ser, mul], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ser, mul = float(tt[0]), float(tt[1])
        ach, ach_m = kb.algo_bytes / (ser / K) / 1e9, kb.algo_bytes / (mul / K) / 1e9
        sfx = "bf16" if dtype_name == "bf16" else "f32"
        return {"rows_per_s": world * B * K / ser, "ms_per_step": 1e3 * ser / K,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "algorithmic_bytes_per_launch": kb.algo_bytes},
                "single_stream_frac": ach / peak,
                "multi_stream": {"streams": kb.nstreams, "ms_per_step": 1e3 * mul / K, "rows_per_s": world * B * K / mul,
                                 "frac": ach_m / peak},
                "kernel": _cabi.describe_energy(B, M, D, sfx), "x0": "fp32 (mixed entry point)" if x0_f32 else sfx,
                "timed_region": f"median of 5 repetitions of exactly {K} single-stream steps, {kb.nsets} rotating sets"}

    other = "bf16" if args.dtype == "f32" else "f32"
    aux[f"k1_{other}"] = k1_block(other)
    aux["k1_bf16_x0f32"] = k1_block("bf16", x0_f32=True)
    torch.cuda.empty_cache()

    # -- the same launch with its inputs L2-RESIDENT (2 rotating sets = 53 MB < L2): what it costs inside a training step, where
    #    the backbone has just written the draws and the backward pass reads the gradient back from L2.  NOT a roofline figure.
    def k1_warm(dtype_name, x0_f32=False):
        kb = K1Bench(L, dev, dtype_name, rank, world, 1, args.no_graph, nsets=2, x0_f32=x0_f32)
        K = 480
        with torch.cuda.stream(kb.stream):
            kb.run_serial(K)
            kb.stream.synchronize()
            ser, _ = kb.timed(kb.run_serial, K, 5)
        return 1e6 * ser / K

    aux["k1_l2_resident"] = {"f32_us_per_launch": k1_warm("f32"), "bf16_us_per_launch": k1_warm("bf16"),
                             "bf16_x0f32_us_per_launch": k1_warm("bf16", x0_f32=True),
                             "note": "one stream, 2 rotating sets (inputs and gradient stay in the 126 MB L2): the cost of the loss "
                                     "launch inside a training step; the headline uses HBM-cold inputs (40 sets)"}
    torch.cuda.empty_cache()

    # -- BASELINE config 3 at its tensor-core point: m = 32, bf16 — the tcgen05 kernel against the blocked packed-fp32 kernel
    def k1_m32(variant):
        _cabi.set_tuning("energy.variant", variant)
        try:
            kb = K1Bench(L, dev, "bf16", rank, world, 4, args.no_graph, m=32)
            K = 240
            with torch.cuda.stream(kb.stream):
                kb.run_serial(K)
                kb.run_multi(K)
                kb.stream.synchronize()
                ser, _ = kb.timed(kb.run_serial, K, 3)
                mul, _ = kb.timed(kb.run_multi, K, 3)
            return {"kernel": _cabi.describe_energy(B, 32, D, "bf16"), "us_per_launch_1_stream": 1e6 * ser / K,
                    "us_per_launch_4_streams": 1e6 * mul / K, "algorithmic_bytes_per_launch": kb.algo_bytes,
                    "frac_hbm_4_streams": kb.algo_bytes / (mul / K) / 1e9 / peak}
        finally:
            _cabi.set_tuning("energy.variant", 0)

    if world == 1:
        tc, blk = k1_m32(7), k1_m32(4)
        aux["k1_m32_bf16"] = {"tensor_core": tc, "blocked_fp32_pipe": blk,
                              "speedup_1_stream": blk["us_per_launch_1_stream"] / tc["us_per_launch_1_stream"],
                              "speedup_4_streams": blk["us_per_launch_4_streams"] / tc["us_per_launch_4_streams"],
                              "note": "BASELINE config 3 (m=32, D=3072, bf16, beta=0.1): Gram + coefficient mixing on tcgen05 "
                                      "(default for m = 32) vs the direct-difference kernel it replaces"}
        torch.cuda.empty_cache()
        aux["single_launch_copy_ceiling"] = copy_ceiling_same_bytes(dev, peak)

    if not args.no_elementwise and world == 1:
        aux["elementwise"] = elementwise_rooflines(dev, peak)

    precisions = [p for p in args.dit_precision.split(",") if p in ("bf16", "tf32", "fp32")]
    if args.dit_steps > 0:
        from ddm_b200 import launcher

        aux["dit_train"] = {"config": "DDDMDiT CIFAR-10 32x32 training step on synthetic images, batch 128/GPU, m=8, "
                                      "data-parallel (one flat-gradient NCCL all-reduce), loss kernels K4+K2c+K1"}
        for prec in precisions:
            steps = args.dit_steps if prec == "bf16" else max(20, args.dit_steps // 2)
            targs = launcher.build_parser().parse_args(["--synthetic", "--precision", prec])
            aux["dit_train"][prec] = launcher.measure_throughput(targs, dev, world, steps=steps, warmup=5)
            torch.cuda.empty_cache()
        first = aux["dit_train"].get(precisions[0]) if precisions else None
        if first:  # the keys earlier rounds' readers look for
            aux["dit_train"].update({k: first[k] for k in ("img_per_s", "ms_per_step", "steps", "precision", "n_gpus")})
        # multi-rank parity on hardware (also run at N = 1: one graph against the eager global step)
        pargs = launcher.build_parser().parse_args(["--synthetic", "--precision", "fp32"])
        aux["dp_parity"] = launcher.dp_parity(pargs, dev, world)
        pargs = launcher.build_parser().parse_args(["--synthetic", "--precision", "bf16"])
        aux["dp_parity_bf16"] = launcher.dp_parity(pargs, dev, world)
        torch.cuda.empty_cache()
    if world > 1:
        # the isolated loss WITH its one data-path collective in the timed region: K4 -> all-reduce(sum_b w) -> K1, eager,
        # one stream.  Dominated by the latency of a 4-byte NCCL all-reduce; in training it hides behind the backbone.
        kb = K1Bench(L, dev, args.dtype, rank, world, 1, True, nsets=8)
        tdev = [torch.rand(B, device=dev) for _ in range(8)]
        wsum = [torch.empty(1, device=dev) for _ in range(8)]

        def step(i):
            s = kb.sets[i % 8]
            _cabi.check(L.dddm_sigmoid_weight_sum_f32(tdev[i % 8].data_ptr(), W_BIAS, None, s["wsum"].data_ptr(), B,
                                                      torch.cuda.current_stream().cuda_stream))
            dist.all_reduce(s["wsum"])
            kb.launch(s, torch.cuda.current_stream().cuda_stream)

        for i in range(10):
            step(i)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 100
        for i in range(n):
            step(i)
        e1.record()
        e1.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        aux["k1_with_weight_allreduce"] = {"ms_per_step": 1e3 * float(tt) / n, "rows_per_s": world * B * n / float(tt),
                                           "steps": n, "launch": "eager: K4, NCCL all-reduce of one float, K1 on one stream"}
        del wsum
    if args.dit_steps > 0 and world == 1 and args.cpu_seconds > 0:
        # CPU baseline of config 4 at a REDUCED batch (a full B=128 CPU step is ~5.7 TFLOP): the
        # reference's step (oracle/torch_port.training_step = dddm/training.py:57-85) + backward + clip + AdamW on the
        # host cores, default DiT in fp32, B=4, m=8.  Reported beside the GPU figure, never on the product path.
        from ddm_b200 import backbones
        from oracle import torch_port

        cb, cm = 4, 8
        torch.manual_seed(0)
        cpu_model = backbones.DDDMDiT()
        cpu_opt = torch.optim.AdamW(cpu_model.parameters(), lr=1e-4, weight_decay=0.01)
        cx0 = torch.rand(cb, 3, 32, 32) * 2 - 1

        def cpu_step():
            t, eps, xi = torch.rand(cb), torch.randn_like(cx0), torch.randn(cb, cm, 3, 32, 32)
            loss = torch_port.training_step(cpu_model, cx0, t, eps, xi, m=cm, beta=BETA, lam=LAM, w_bias=W_BIAS)[0]
            cpu_opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(cpu_model.parameters(), 1.0)
            cpu_opt.step()

        cpu_step()
        c0 = time.perf_counter()
        nsteps = 3
        for _ in range(nsteps):
            cpu_step()
        cdt = (time.perf_counter() - c0) / nsteps
        aux["dit_train"]["cpu_baseline"] = {"img_per_s": cb / cdt, "s_per_step": cdt, "batch": cb, "m": cm, "kind": "port",
                                            "cores": torch.get_num_threads(),
                                            "sample": f"{nsteps} steps at the reduced batch {cb} (fp32, eager PyTorch)"}
    if args.sampler_samples > 0:
        from ddm_b200.backbones import DDDMDiT
        from ddm_b200.sampling import sample_dddm

        per = max(1, args.sampler_samples // world)
        net = DDDMDiT().to(dev)
        if "bf16" in precisions or not precisions:
            net = net.to(torch.bfloat16)  # bf16 weights and activations; x_t, the noise and the K3 update stay fp32
        res = {}
        if True:
            for nsteps in (20, 100):
                sample_dddm(net, per, steps=2, device=str(dev), data_shape=(3, 32, 32), cuda_graph=True)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                sample_dddm(net, per, steps=nsteps, device=str(dev), data_shape=(3, 32, 32), cuda_graph=True)
                a1.record()
                a1.synchronize()
                dt = a0.elapsed_time(a1) * 1e-3
                if world > 1:
                    tt = torch.tensor([dt], device=dev, dtype=torch.float64)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    dt = float(tt)
                res[f"steps{nsteps}"] = {"samples_per_s": per * world / dt, "seconds": dt}
        aux["sampler"] = {"n_samples": per * world, "per_gpu": per, "model": "DDDMDiT(default)",
                          "launch": "one cached CUDA graph per Algorithm-2 step (captured by the warm-up call)", **res}
    if args.mmd_samples > 0 and world == 1:
        import ddm_b200

        n = args.mmd_samples
        gen = torch.Generator(device=dev).manual_seed(0)
        xs = torch.rand(n, D, device=dev, generator=gen) * 2 - 1
        ys = torch.rand(n, D, device=dev, generator=gen) * 1.9 - 0.95

        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            for _ in range(reps):
                out = fn()
            m1.record()
            m1.synchronize()
            return m0.elapsed_time(m1) / reps, float(out)

        torch.backends.cuda.matmul.allow_tf32 = False  # fp32 Gram, as the reference computes it
        ours_ms, ours_v = timed(lambda: ddm_b200.rbf_mmd2(xs, ys, 45.0))
        tf32_ms, tf32_v = timed(lambda: ddm_b200.rbf_mmd2(xs, ys, 45.0, allow_tf32=True))

        def eager():  # the reference's formula (dddm/metrics.py:140-163) in eager PyTorch on the same GPU
            def pd(a, b):
                return (a * a).sum(-1).unsqueeze(-1) + (b * b).sum(-1).unsqueeze(0) - 2.0 * (a @ b.T)
            g = 1.0 / (2.0 * 45.0**2)
            mask = ~torch.eye(n, dtype=torch.bool, device=dev)
            return torch.exp(-g * pd(xs, xs))[mask].mean() + torch.exp(-g * pd(ys, ys))[mask].mean() - 2.0 * torch.exp(
                -g * pd(xs, ys)).mean()

        ref_ms, ref_v = timed(eager)
        aux["rbf_mmd2"] = {"n": n, "m": n, "D": D, "ms": ours_ms, "value": ours_v, "tf32_gram_ms": tf32_ms,
                           "tf32_gram_value": tf32_v, "eager_reference_formula_ms": ref_ms,
                           "eager_value": ref_v, "note": "evaluation-side pairwise kernel (SURVEY 8f-4). ms: the default path at this size, ONE fused tcgen05 "
                           "kernel per term (bf16 hi/lo split of the fp32 inputs, Gram tile in tensor memory, distance / exp / "
                           "mask / sum epilogue, no n x n matrix in HBM; csrc/metrics_tc.cu). tf32_gram_ms: the library-GEMM tile "
                           "path with allow_tf32=True, for comparison"}
    return aux


class K1Bench:
    """The isolated fused loss on rotating HBM-cold buffer sets, timed two ways: strictly serialized on ONE stream
    (the headline: what a training step sees, one loss launch between the backbone's forward and backward) and with
    the independent steps issued round-robin on several streams (aggregate: consecutive minibatches overlap)."""

    def __init__(self, L, dev, dtype_name, rank, world, nstreams, no_graph, nsets=0, x0_f32=False, m=0):
        import torch
        import torch.distributed as dist

        from ddm_b200 import _cabi

        M = m or globals()["M"]  # BASELINE config 3 runs the same harness at m = 32
        self.m = M
        self.torch, self.L, self.dev, self.world = torch, L, dev, world
        self.dtype_name = dtype_name
        tdtype = torch.float32 if dtype_name == "f32" else torch.bfloat16
        esz = 4 if dtype_name == "f32" else 2
        self.algo_bytes = (2 * B * M * D) * esz + B * D * (4 if x0_f32 else esz)  # read xhat + x0, write grad
        self.nsets = nsets or max(4, -(-8 * L2_BYTES // self.algo_bytes))  # working set > 8x L2: every launch reads HBM-cold data
        # a set (and its workspace) is always launched on the same stream of the multi-stream runner: the C-ABI contract is one
        # workspace per stream (concurrent launches sharing one would mix their row sums)
        self.nsets = -(-self.nsets // max(1, nstreams)) * max(1, nstreams)
        fn = getattr(L, f"dddm_energy_fused_{dtype_name}" + ("_x0f32" if x0_f32 else ""))
        self.sets = []
        for s in range(self.nsets):
            xh, x0, t = make_inputs(1000 * rank + s, tdtype, M)
            if x0_f32:
                x0 = make_inputs(1000 * rank + s, torch.float32, M)[1]
            self.sets.append({"xh": xh.to(dev), "x0": x0.to(dev), "t": t.to(dev),
                              "grad": torch.empty(B, M, D, dtype=tdtype, device=dev), "out": torch.zeros(4, device=dev),
                              "wsum": torch.empty(1, device=dev),
                              "ws": torch.zeros(L.dddm_energy_workspace_bytes(B, M), dtype=torch.uint8, device=dev)})
        self.stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(self.stream):
            for s in self.sets:  # W = mean_b w(t_b): an input of the isolated loss, computed once outside the timed region
                _cabi.check(L.dddm_sigmoid_weight_sum_f32(s["t"].data_ptr(), W_BIAS, None, s["wsum"].data_ptr(), B,
                                                          self.stream.cuda_stream))
            if world > 1:  # global-batch weight: one float all-reduce, outside the timed region
                for s in self.sets:
                    dist.all_reduce(s["wsum"])
        self.stream.synchronize()
        wscale = 1.0 / (B * world)

        def launch(s, cuda_stream):
            _cabi.check(fn(s["xh"].data_ptr(), s["x0"].data_ptr(), s["wsum"].data_ptr(), wscale, s["grad"].data_ptr(),
                           s["out"].data_ptr(), s["ws"].data_ptr(), B, M, D, BETA, LAM, cuda_stream))

        self.launch = launch
        self.no_graph = no_graph
        self.nstreams = 1 if no_graph else max(1, nstreams)
        self.run_serial = self._runner(1)
        self.run_multi = self._runner(self.nstreams) if self.nstreams > 1 else self.run_serial

    def _runner(self, nstreams):
        torch, dev, stream, sets, nsets = self.torch, self.dev, self.stream, self.sets, self.nsets
        sides = [torch.cuda.Stream(dev) for _ in range(nstreams - 1)]

        def capture(n):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                main = torch.cuda.current_stream()
                for sd in sides:
                    sd.wait_stream(main)
                for i in range(n):
                    st = main if i % nstreams == 0 else sides[i % nstreams - 1]
                    self.launch(sets[i % nsets], st.cuda_stream)
                for sd in sides:
                    main.wait_stream(sd)
            return g

        if self.no_graph:
            def run_plain(n):
                for i in range(n):
                    self.launch(sets[i % nsets], stream.cuda_stream)
            return run_plain
        chunk = nsets * max(1, 480 // nsets)
        graphs = {}

        def run_graph(n):
            full, rem = divmod(n, chunk)
            for size, count in ((chunk, full), (rem, 1 if rem else 0)):
                if count:
                    if size not in graphs:
                        graphs[size] = capture(size)
                    for _ in range(count):
                        graphs[size].replay()
        return run_graph

    def timed(self, runner, K, reps):
        """Median over `reps` repetitions of EXACTLY K steps between two CUDA events on the launch stream."""
        torch = self.torch
        times = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            runner(K)
            e1.record(self.stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1) * 1e-3)
        return sorted(times)[len(times) // 2], times


def copy_ceiling(dev, h2d_bytes, d2h_bytes, reps=20):
    """Plain pinned-memory copies of one step's bytes, no kernels: H2D alone, D2H alone, both directions at once (two
    streams).  The ceiling the host-buffer e2e figure is measured against — a number for "PCIe/host limited"."""
    import torch

    hin, hout = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory(), torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    din, dout = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev), torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(up, down):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        return (time.perf_counter() - t0) / reps

    run(True, True)
    run(True, True)
    # a ceiling: the best of three passes per direction (the first passes over freshly pinned pages run 5-10 % slower)
    t_up, t_down, t_both = (min(run(*ud) for _ in range(3)) for ud in ((True, False), (False, True), (True, True)))
    return {"h2d_gbs": h2d_bytes / t_up / 1e9, "d2h_gbs": d2h_bytes / t_down / 1e9,
            "duplex_h2d_gbs": h2d_bytes / t_both / 1e9, "duplex_d2h_gbs": d2h_bytes / t_both / 1e9,
            "duplex_s_per_step": t_both, "bytes": [h2d_bytes, d2h_bytes],
            "how": f"{reps} back-to-back cudaMemcpyAsync of one step's bytes per direction from/to pinned host memory, "
                   "two streams, wall clock around a device synchronize; best of three passes after two warm-up passes"}


def copy_ceiling_same_bytes(dev, peak):
    """What ONE dependent launch can reach at K1's size: the driver's own device-to-device copy of the gradient's bytes
    (12.6 MB read + 12.6 MB written), back to back on one stream over rotating HBM-cold buffers — no arithmetic, no
    row-wise dependency.  The kernel's single-stream figure is to be read against this, not against the burst peak."""
    import torch

    nbytes = B * M * D * 4
    nsets = max(4, -(-8 * L2_BYTES // (2 * nbytes)))
    src = [torch.empty(nbytes, dtype=torch.uint8, device=dev).fill_(i & 255) for i in range(nsets)]
    dst = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(nsets)]
    stream = torch.cuda.Stream(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        for i in range(3):
            dst[i].copy_(src[i])
    stream.synchronize()
    with torch.cuda.graph(g, stream=stream):
        for i in range(nsets):
            dst[i].copy_(src[i])
    ts = []
    with torch.cuda.stream(stream):
        g.replay()
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                g.replay()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3 / (5 * nsets))
    sec = sorted(ts)[2]
    k1_bytes = (2 * B * M * D + B * D) * 4
    return {"what": "cudaMemcpyAsync device-to-device of B*m*D*4 bytes, one stream, back to back, rotating sets",
            "bytes_moved": 2 * nbytes, "us_per_copy": sec * 1e6, "GBps": 2 * nbytes / sec / 1e9,
            "frac_of_peak": 2 * nbytes / sec / 1e9 / peak,
            "k1_bytes_at_this_rate_us": k1_bytes / (2 * nbytes / sec) * 1e6}


def elementwise_rooflines(dev, peak):
    """K2 / K2c / K3 at their BASELINE shapes, same discipline as K1: rotating buffers beyond 8x L2, one stream, CUDA
    graph of back-to-back launches, CUDA events."""
    import torch

    from ddm_b200 import ops

    res = {}
    C, H, Wd = 3, 32, 32

    def bench(name, make, call, nbytes, nstreams=4):
        # every launch of the graph has its own inputs AND its own outputs (the ops allocate their results: they are kept
        # alive until the graph is dropped, otherwise the allocator would hand every launch the same, L2-resident block)
        nsets = max(4, -(-8 * L2_BYTES // nbytes))
        nsets = -(-nsets // nstreams) * nstreams
        sets = [make(i) for i in range(nsets)]
        stream = torch.cuda.Stream(dev)
        side = [torch.cuda.Stream(dev) for _ in range(nstreams - 1)]
        with torch.cuda.stream(stream):
            for s_ in sets[:3]:
                call(s_)
        stream.synchronize()

        def capture(lanes):
            g, keep = torch.cuda.CUDAGraph(), []
            with torch.cuda.graph(g, stream=stream):
                if len(lanes) > 1:
                    fork = torch.cuda.Event()
                    fork.record(stream)
                    for l_ in lanes[1:]:
                        l_.wait_event(fork)
                for i, s_ in enumerate(sets):
                    with torch.cuda.stream(lanes[i % len(lanes)]):
                        keep.append(call(s_))
                for l_ in lanes[1:]:
                    join = torch.cuda.Event()
                    join.record(l_)
                    stream.wait_event(join)
            return g, keep

        def timed(g):
            ts = []
            with torch.cuda.stream(stream):
                g.replay()
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    g.replay()
                    e1.record(stream)
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e-3 / nsets)
            return sorted(ts)[2]

        g, keep = capture([stream])
        sec = timed(g)
        res[name] = {"us_per_launch": sec * 1e6, "algorithmic_bytes": nbytes,
                     "roofline": {"bound": "hbm", "achieved": nbytes / sec / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": nbytes / sec / 1e9 / peak}, "rotating_sets": nsets,
                     "buffers": "inputs and outputs distinct per launch (HBM-cold)"}
        del g, keep
        torch.cuda.empty_cache()
        try:  # labelled AGGREGATE: independent launches round-robin on several streams inside one graph
            g, keep = capture([stream] + side)
            sec_n = timed(g)
            res[name]["multi_stream"] = {"streams": nstreams, "us_per_launch": sec_n * 1e6,
                                         "frac": nbytes / sec_n / 1e9 / peak,
                                         "note": "AGGREGATE with independent launches in flight, not the per-launch figure"}
            del g, keep
        except Exception as exc:  # never lose the single-stream figure to the auxiliary one
            res[name]["multi_stream"] = {"error": repr(exc)[:200]}
        del sets
        torch.cuda.empty_cache()

    bench("K2_forward_marginal_expand_f32",
          lambda i: (torch.rand(B, C, H, Wd, device=dev), torch.rand(B, device=dev), torch.randn(B, C, H, Wd, device=dev)),
          lambda s_: ops.forward_marginal_expand(s_[0], s_[1], s_[2], M, False), (2 * B * D + B * M * D) * 4 + 4 * B)
    bench("K2c_forward_marginal_concat_f32_to_bf16",
          lambda i: (torch.rand(B, C, H, Wd, device=dev), torch.rand(B, device=dev), torch.randn(B, C, H, Wd, device=dev),
                     torch.randn(B, M, C, H, Wd, device=dev)),
          lambda s_: ops.forward_marginal_concat(s_[0], s_[1], s_[2], s_[3], True, 4),
          (2 * B * D + B * M * D) * 4 + 2 * B * M * D * 2 + B * D * 4)
    n = 1024
    bench("K3_bridge_step_f32_1024",
          lambda i: (torch.randn(n, C, H, Wd, device=dev), torch.randn(n, C, H, Wd, device=dev),
                     torch.randn(n, C, H, Wd, device=dev), torch.tensor([0.45], device=dev), torch.tensor([0.5], device=dev)),
          lambda s_: ops.bridge_step(s_[0], s_[1], s_[2], s_[3], s_[4], 1.0), 4 * n * D * 4)
    # K3 with the step's two Gaussian draws fused in (z in registers, next xi written): replaces randn + randn + K3
    bench("K3_bridge_step_fused_noise_f32_1024",
          lambda i: (torch.randn(n, C, H, Wd, device=dev), torch.randn(n, C, H, Wd, device=dev),
                     torch.empty(n, C, H, Wd, device=dev), torch.tensor([0.45], device=dev), torch.tensor([0.5], device=dev)),
          lambda s_: ops.bridge_step_philox_(s_[0], s_[1], s_[2], s_[3], s_[4], 1.0, seed=1234, offset_z=4096, offset_xi=8192),
          4 * n * D * 4)
    bench("unfused_randn_randn_K3_f32_1024",
          lambda i: (torch.randn(n, C, H, Wd, device=dev), torch.randn(n, C, H, Wd, device=dev),
                     torch.tensor([0.45], device=dev), torch.tensor([0.5], device=dev)),
          lambda s_: (torch.randn_like(s_[0]), ops.bridge_step(s_[0], s_[1], torch.randn_like(s_[0]), s_[2], s_[3], 1.0)),
          6 * n * D * 4)
    return res


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget (rank 0, N=1 only)")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of CUDA graphs")
    ap.add_argument("--sets", type=int, default=0, help="rotating input sets (0 = enough to exceed 8x L2)")
    ap.add_argument("--streams", type=int, default=6,
                    help="streams of the AGGREGATE figure (independent steps round-robin inside one CUDA graph); the headline "
                         "value is always the single-stream one")
    ap.add_argument("--tune", default="", help="comma list key=value for dddm_set_tuning, e.g. energy.cluster=4")
    ap.add_argument("--dit-steps", type=int, default=100,
                    help="auxiliary: DP DiT training steps to time for the img/s figure (0 = skip)")
    ap.add_argument("--dit-precision", default="bf16,tf32,fp32", help="comma list of backbone precisions to time")
    ap.add_argument("--sampler-samples", type=int, default=1024, help="auxiliary: Algorithm-2 samples (0 = skip)")
    ap.add_argument("--mmd-samples", type=int, default=10000, help="auxiliary: rbf_mmd2 set size (0 = skip)")
    ap.add_argument("--no-elementwise", action="store_true", help="skip the K2/K2c/K3 rooflines")
    ap.add_argument("--aux-timeout", type=float, default=420.0, help="seconds the auxiliary measurements may take")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from ddm_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _cabi.lib()
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        _cabi.set_tuning(k, int(v))

    tdtype = torch.float32 if args.dtype == "f32" else torch.bfloat16
    esz = 4 if args.dtype == "f32" else 2
    K, W = max(1, args.steps), max(3, args.warmup)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, burst)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    kb = K1Bench(L, dev, args.dtype, rank, world, args.streams, args.no_graph, nsets=args.sets)
    algo_bytes, nsets, stream = kb.algo_bytes, kb.nsets, kb.stream
    launches_before = _cabi.launch_count()
    with torch.cuda.stream(stream):
        kb.run_serial(W)
        kb.run_serial(K)  # extra untimed pass of the whole region: clocks and caches in steady state (also captures graphs)
        kb.run_multi(min(K, 2000))
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        # HEADLINE: exactly K steps strictly serialized on ONE stream; repeated a few times (median) because a ~9 us
        # kernel makes a single short region noisy
        reps = 5 if K * 1e-5 < 0.5 else 1
        elapsed, times = kb.timed(kb.run_serial, K, reps)
        # aggregate: the same steps round-robin on several streams (labelled as such, not the headline)
        km = min(K, 2000)
        multi_elapsed, _ = kb.timed(kb.run_multi, km, 3)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.stop()
    if world > 1:
        tmax = torch.tensor([elapsed, multi_elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed, multi_elapsed = float(tmax[0]), float(tmax[1])
    timed_launches = _cabi.launch_count() - launches_before
    ms_per_step = 1e3 * elapsed / K
    multi_ms = 1e3 * multi_elapsed / km
    value = world * B * K / elapsed
    del kb.sets[:]  # free the 1 GiB of rotating sets before the other measurements
    kernel_desc = _cabi.describe_energy(B, M, D, args.dtype)

    # ---- e2e: C-ABI host-buffer session (pinned host inputs in the session's packed layout: ONE H2D copy of
    #      [xhat | x0 | t], K4 + K1, ONE D2H copy of [grad | out] per step; 4-deep pipeline)
    e2e_steps = max(3, min(args.e2e_steps, K))
    sess = L.dddm_session_create(B, M, D, 0 if args.dtype == "f32" else 1, local_rank)
    if not sess:
        raise SystemExit(f"dddm_session_create failed: {_cabi.strerror(L.dddm_last_error())}")
    import ctypes

    sz = [ctypes.c_size_t() for _ in range(5)]
    _cabi.check(L.dddm_session_packed_layout(sess, *[ctypes.addressof(v) for v in sz]))
    in_bytes, x0_off, t_off, out_bytes, out_off = [int(v.value) for v in sz]
    host = []
    for s in range(4):
        xh, x0, t = make_inputs(5000 + 10 * rank + s, tdtype)
        pin, pout = torch.zeros(in_bytes, dtype=torch.uint8).pin_memory(), torch.zeros(out_bytes, dtype=torch.uint8).pin_memory()
        pin[:xh.numel() * esz] = xh.contiguous().view(torch.uint8).reshape(-1)
        pin[x0_off:x0_off + x0.numel() * esz] = x0.contiguous().view(torch.uint8).reshape(-1)
        pin[t_off:t_off + 4 * B] = t.contiguous().view(torch.uint8).reshape(-1)
        host.append((pin, pout))

    def e2e_run(n):
        for i in range(n):
            pin, pout = host[i % 4]
            _cabi.check(L.dddm_session_enqueue_host(sess, pin.data_ptr(), pin.data_ptr() + x0_off, pin.data_ptr() + t_off,
                                                    W_BIAS, BETA, LAM, pout.data_ptr(), pout.data_ptr() + out_off))
        _cabi.check(L.dddm_session_wait(sess))

    e2e_run(8)
    e2e_run(e2e_steps)  # one untimed pass of the whole region, as for the kernel: host page tables / IOMMU entries of every buffer warm
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_all = []
    for _ in range(5):  # median of five repetitions of exactly e2e_steps steps (a 6 ms region is noisy on a shared host)
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        e2e_all.append(time.perf_counter() - t0)
        if world > 1:
            dist.barrier()
    e2e_dt = sorted(e2e_all)[len(e2e_all) // 2]
    if world > 1:
        tmax = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_dt = float(tmax)
    L.dddm_session_destroy(sess)
    e2e_value = world * B * e2e_steps / e2e_dt
    h2d = t_off + 4 * B          # bytes of the one upload (payload + alignment padding of the packed layout)
    d2h = out_off + 16           # bytes of the one download
    if world > 1:
        dist.barrier()
    ceiling = copy_ceiling(dev, h2d, d2h)  # all ranks at once: the host's aggregate copy rate is the limiter at N = 8
    if world > 1:
        tmax = torch.tensor([ceiling["duplex_s_per_step"]], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ceiling["duplex_s_per_step"] = float(tmax)
    ceiling["rows_per_s_at_ceiling"] = world * B / ceiling["duplex_s_per_step"]
    del host

    achieved = algo_bytes / (elapsed / K) / 1e9
    multi_achieved = algo_bytes / (multi_elapsed / km) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", f"k1_traffic_{args.dtype}.json")
    if os.path.exists(tpath):  # dram bytes of ONE launch from the committed ncu --set full capture (tools/ncu_summary.py)
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")

    cpu = eager = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        eager = gpu_eager_rate(dev)
        rate, n, threads, mean, best = cpu_port_rate(args.cpu_seconds, args.dtype)
        cpu = {"value": rate, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": f"{n} full fwd+bwd passes over the same B={B},m={M},D={D} fp32 batch with oracle/torch_port.py "
                         f"(eager PyTorch + autograd, the reference's CPU path), mean {mean * 1e3:.1f} ms, best "
                         f"{best * 1e3:.1f} ms, os.cpu_count()={os.cpu_count()}"}
    line = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": workload_name(args.dtype), "rows_per_gpu": B, "global_rows": B * world,
                   "parallelism": f"dp{world} (independent row shards, global weight pre-reduced)",
                   "l2_policy": f"{nsets} rotating input/output sets = {nsets * algo_bytes / 2**20:.0f} MiB > 8x L2 "
                                f"(every launch reads HBM-cold inputs)",
                   "launch": ("ONE stream, launches strictly serialized (CUDA graph of back-to-back launches, programmatic "
                              "dependent launch): the per-launch figure a training step sees" if not args.no_graph
                              else "python launches, one stream"),
                   "single_stream_ms_per_step": ms_per_step,
                   "single_stream_rows_per_s": B / (ms_per_step * 1e-3),
                   "multi_stream": {"streams": kb.nstreams, "ms_per_step": multi_ms, "rows_per_s": world * B / (multi_ms * 1e-3),
                                    "note": "AGGREGATE, not the headline: independent minibatches issued round-robin on "
                                            "several streams inside one CUDA graph so that consecutive launches overlap"},
                   "timed_region": f"median of {len(times)} repetitions of exactly {K} steps (CUDA events on the "
                                   f"launch stream)", "kernel": kernel_desc, "tuning": args.tune or "auto"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                     "frac_of_8TBs_nominal": achieved / 8000.0,
                     "single_stream_frac": achieved / peak,
                     "multi_stream_frac": multi_achieved / peak, "multi_stream_achieved": multi_achieved,
                     "note": "achieved / frac are the single-launch (one stream) figures; multi_stream_* is the aggregate "
                             "with several launches in flight; profiles/r02_k1_single_launch.md explains the gap"},
        "cpu_baseline": cpu,
        "gpu_eager_baseline": eager,
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / e2e_steps,
                "all_region_ms_per_step": [1e3 * t / e2e_steps for t in e2e_all],
                "path": "dddm_session_enqueue_host (C ABI; pinned host buffers in the session's packed layout: one "
                        "cudaMemcpyAsync per direction and step; 4-deep pipeline) + dddm_session_wait",
                "host_copy_ceiling": ceiling,
                "frac_of_copy_ceiling": e2e_value / ceiling["rows_per_s_at_ceiling"],
                "cpu_affinity": affinity},
        "gpu_launches": K,
        "gpu_launches_counted": {"headline_region": K, "library_launches_since_warmup_start": int(timed_launches)},
        "clocks": sampler.summary(),
        "all_region_ms_per_step": [1e3 * x / K for x in times],
    }

    # The auxiliary figures must never cost the headline line: a watchdog prints the line without them and ends
    # every rank if they do not finish in time (e.g. a collective that never completes).
    emitted = threading.Lock()

    def emit(aux_value) -> None:
        if emitted.acquire(blocking=False) and rank == 0:
            line["aux"] = aux_value
            print(json.dumps(line), flush=True)

    def on_timeout() -> None:
        emit({"error": f"auxiliary measurements did not finish within {args.aux_timeout:.0f} s and were dropped",
              "partial": AUX_PARTIAL})
        os._exit(0)

    watchdog = threading.Timer(args.aux_timeout, on_timeout)
    watchdog.daemon = True
    watchdog.start()
    try:
        aux = run_aux(args, dev, world, rank, L, peak)
    except Exception as exc:  # noqa: BLE001 - report, keep the headline
        aux = {"error": f"{type(exc).__name__}: {exc}", "partial": AUX_PARTIAL}
    watchdog.cancel()
    emit(aux)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
