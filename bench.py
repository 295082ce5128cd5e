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


def make_inputs(seed: int, dtype):
    """SURVEY.md §8(d) 'late' regime (the precision-critical one): x0 in CIFAR range, draws x0 + 0.05 noise."""
    import torch

    gen = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, D, generator=gen).clamp(-1, 1)
    xh = x0[:, None, :] + 0.05 * torch.randn(B, M, D, generator=gen)
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
    warmup = max(1, min(args.warmup, 3))
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
        "config": {"workload": workload_name("f32"), "device": "host CPU", "threads": torch.get_num_threads()},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{steps} full fwd+bwd passes over the B={B} batch (oracle/torch_port.py, eager "
                                   f"PyTorch + autograd as in the reference), os.cpu_count()={os.cpu_count()}"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_aux(args, dev, world) -> dict:
    """Auxiliary, outside the timed region: BASELINE configs 4 (DP DiT training img/s) and 5 (Algorithm-2 sampler)."""
    import torch
    import torch.distributed as dist

    aux = {}
    if args.dit_steps > 0:
        from ddm_b200 import launcher

        targs = launcher.build_parser().parse_args(["--synthetic", "--precision", args.dit_precision])
        aux["dit_train"] = launcher.measure_throughput(targs, dev, world, steps=args.dit_steps, warmup=3)
        aux["dit_train"]["config"] = ("DDDMDiT CIFAR-10 32x32 training step on synthetic images, batch 128/GPU, m=8, "
                                      "data-parallel (one flat-gradient NCCL all-reduce), loss kernels K4+K2c+K1")
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
        if args.dit_precision == "bf16":
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
                           "eager_value": ref_v, "note": "evaluation-side pairwise kernel (SURVEY 8f-4): fp32 GEMM tiles "
                           "(cuBLAS; only the upper trapezoids of the symmetric xx/yy terms) + one fused "
                           "distance/exp/mask/sum pass per tile"}
    return aux


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
                    help="streams the independent steps are issued on round-robin inside the CUDA graph (1 = serialized)")
    ap.add_argument("--tune", default="", help="comma list key=value for dddm_set_tuning, e.g. energy.cluster=4")
    ap.add_argument("--dit-steps", type=int, default=10,
                    help="auxiliary: DP DiT training steps to time for the img/s figure (0 = skip)")
    ap.add_argument("--dit-precision", default="bf16", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--sampler-samples", type=int, default=1024, help="auxiliary: Algorithm-2 samples (0 = skip)")
    ap.add_argument("--mmd-samples", type=int, default=10000, help="auxiliary: rbf_mmd2 set size (0 = skip)")
    ap.add_argument("--aux-timeout", type=float, default=240.0, help="seconds the auxiliary measurements may take")
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
    algo_bytes = (2 * B * M * D + B * D) * esz  # read xhat + x0, write grad
    nsets = args.sets or max(4, -(-8 * L2_BYTES // algo_bytes))  # working set > 8x L2: every launch reads HBM-cold data
    fn = getattr(L, f"dddm_energy_fused_{args.dtype}")
    K, W = max(1, args.steps), max(3, args.warmup)

    sets = []
    for s in range(nsets):
        xh, x0, t = make_inputs(1000 * rank + s, tdtype)
        sets.append({"xh": xh.to(dev), "x0": x0.to(dev), "t": t.to(dev), "grad": torch.empty(B, M, D, dtype=tdtype, device=dev),
                     "out": torch.zeros(4, device=dev), "wsum": torch.empty(1, device=dev),
                     "ws": torch.zeros(L.dddm_energy_workspace_bytes(B, M), dtype=torch.uint8, device=dev)})
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        for s in sets:  # W = mean_b w(t_b): an input of the isolated loss, computed once outside the timed region
            _cabi.check(L.dddm_sigmoid_weight_sum_f32(s["t"].data_ptr(), W_BIAS, None, s["wsum"].data_ptr(), B,
                                                      stream.cuda_stream))
        if world > 1:  # global-batch weight: one float all-reduce, outside the timed region
            for s in sets:
                dist.all_reduce(s["wsum"])
    stream.synchronize()
    wscale = 1.0 / (B * world)

    def launch(s, cuda_stream):
        _cabi.check(fn(s["xh"].data_ptr(), s["x0"].data_ptr(), s["wsum"].data_ptr(), wscale, s["grad"].data_ptr(),
                       s["out"].data_ptr(), s["ws"].data_ptr(), B, M, D, BETA, LAM, cuda_stream))

    def make_runner(nstreams):
        """Returns run_steps(n): n fused launches over the rotating sets.  With nstreams > 1 the steps
        (independent minibatches) are issued round-robin on several streams forked/joined inside the graph."""
        sides = [torch.cuda.Stream(dev) for _ in range(nstreams - 1)]

        def capture(n):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                main = torch.cuda.current_stream()
                for sd in sides:
                    sd.wait_stream(main)
                for i in range(n):
                    st = main if i % nstreams == 0 else sides[i % nstreams - 1]
                    launch(sets[i % nsets], st.cuda_stream)
                for sd in sides:
                    main.wait_stream(sd)
            return g

        if args.no_graph:
            def run_plain(n):
                for i in range(n):
                    launch(sets[i % nsets], stream.cuda_stream)
            return run_plain
        chunk = nsets * max(1, 480 // nsets)
        graphs = {chunk: capture(chunk)}

        def run_graph(n):
            full, rem = divmod(n, chunk)
            for _ in range(full):
                graphs[chunk].replay()
            if rem:
                if rem not in graphs:
                    graphs[rem] = capture(rem)
                graphs[rem].replay()
        return run_graph

    use_graph = not args.no_graph
    nstreams = max(1, args.streams) if use_graph else 1
    run_steps = make_runner(nstreams)
    run_serial = make_runner(1) if nstreams > 1 else run_steps

    launches_before = _cabi.launch_count()
    with torch.cuda.stream(stream):
        run_steps(W)
        run_steps(K)  # extra untimed pass of the whole region: clocks and caches in steady state (also captures graphs)
        run_serial(min(K, 2000))
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        # repeat the K-step timed region a few times and keep the median: a 5 us kernel makes a single
        # short region noisy; every repetition times EXACTLY K steps between two events on the launch stream
        reps = 5 if K * 5e-6 < 0.5 else 1
        times = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run_steps(K)
            e1.record(stream)
            e1.synchronize()
            times.append(e0.elapsed_time(e1) * 1e-3)
        # the same steps strictly serialized on ONE stream (reported for transparency, not the headline)
        ks = min(K, 2000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run_serial(ks)
        e1.record(stream)
        e1.synchronize()
        serial_ms = e0.elapsed_time(e1) / ks
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.stop()
    elapsed = sorted(times)[len(times) // 2]
    if world > 1:
        tmax = torch.tensor([elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed = float(tmax)
    del launches_before
    ms_per_step = 1e3 * elapsed / K
    value = world * B * K / elapsed

    # ---- e2e: C-ABI host-buffer session (pinned host inputs, H2D + K4 + K1 + D2H of loss & grad every step)
    e2e_steps = max(3, min(args.e2e_steps, K))
    host = []
    for s in range(3):
        xh, x0, t = make_inputs(5000 + 10 * rank + s, tdtype)
        host.append((xh.contiguous().pin_memory(), x0.contiguous().pin_memory(), t.contiguous().pin_memory(),
                     torch.empty(B, M, D, dtype=tdtype).pin_memory(), torch.zeros(4).pin_memory()))
    sess = L.dddm_session_create(B, M, D, 0 if args.dtype == "f32" else 1, local_rank)
    if not sess:
        raise SystemExit(f"dddm_session_create failed: {_cabi.strerror(L.dddm_last_error())}")

    def e2e_run(n):
        for i in range(n):
            xh, x0, t, g, o = host[i % 3]
            _cabi.check(L.dddm_session_enqueue_host(sess, xh.data_ptr(), x0.data_ptr(), t.data_ptr(), W_BIAS, BETA, LAM,
                                                    g.data_ptr(), o.data_ptr()))
        _cabi.check(L.dddm_session_wait(sess))

    e2e_run(5)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_dt = float(tmax)
    L.dddm_session_destroy(sess)
    e2e_value = world * B * e2e_steps / e2e_dt
    h2d = (B * M * D + B * D) * esz + B * 4
    d2h = B * M * D * esz + 16

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, burst)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    achieved = algo_bytes / (elapsed / K) / 1e9

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
                   "launch": (f"CUDA graphs, programmatic dependent launch; the independent steps are issued round-robin "
                              f"on {nstreams} streams (fork/join inside the graph) so consecutive minibatches overlap"
                              if use_graph else "python launches, one stream"),
                   "single_stream_ms_per_step": serial_ms,
                   "single_stream_rows_per_s": B / (serial_ms * 1e-3),
                   "timed_region": f"median of {len(times)} repetitions of exactly {K} steps (CUDA events on the "
                                   f"launch stream)", "kernel": _cabi.describe_energy(B, M, D, args.dtype),
                   "tuning": args.tune or "auto"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src,
                     "frac_of_8TBs_nominal": achieved / 8000.0,
                     "single_stream_frac": algo_bytes / (serial_ms * 1e-3) / 1e9 / peak},
        "cpu_baseline": cpu,
        "gpu_eager_baseline": eager,
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / e2e_steps,
                "path": "dddm_session_enqueue_host (C ABI, pinned host buffers, 3-deep pipeline) + dddm_session_wait",
                "cpu_affinity": affinity},
        "gpu_launches": K,
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
        emit({"error": f"auxiliary DiT/sampler measurements did not finish within {args.aux_timeout:.0f} s and were dropped"})
        os._exit(0)

    watchdog = threading.Timer(args.aux_timeout, on_timeout)
    watchdog.daemon = True
    watchdog.start()
    try:
        aux = run_aux(args, dev, world)
    except Exception as exc:  # noqa: BLE001 - report, keep the headline
        aux = {"error": f"{type(exc).__name__}: {exc}"}
    watchdog.cancel()
    emit(aux)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
