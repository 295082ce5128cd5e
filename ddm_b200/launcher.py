"""Data-parallel DDDM training on one box of B200s: one process per GPU, NCCL over NVLink.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m ddm_b200.launcher --synthetic ...

The reference's ``train_cifar10_dit.py`` is single-process, single-device (``:99-100``); this is the
data-parallel launcher BASELINE.json asks for, with the same flag names and defaults
(``train_cifar10_dit.py:362-398``), the same per-step order (``:152-169``: step -> zero_grad ->
backward -> clip_grad_norm_ -> AdamW.step) and the same checkpoint payload (``:32-37``).

Per step and per rank:
  K4  w-sum of the local t  ->  1-float NCCL all-reduce (async, overlaps the backbone forward)
  K2  x_t = alpha x0 + sigma eps written m-fold straight into the backbone input
  backbone forward (PyTorch; bf16 shadow weights + bf16 activations by default, fp32 master weights)
  K1  fused energy-score loss forward + backward with the GLOBAL weight
  backbone backward into ONE flat gradient buffer -> one NCCL all-reduce(AVG) over NVLink/NVSwitch
  device-side global-norm clip + fused AdamW; metrics stay on the device and are read back packed,
  once per log interval.  The whole step replays as one CUDA graph (--cuda-graph).
The batch is sharded by row (``--batch`` is per GPU, BASELINE config 4: 128/GPU); there is no
other parallelism axis (14.5 M parameters, 64-token sequences).
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import time

import torch
import torch.distributed as dist

from .backbones import DDDMDiT
from .sampling import sample_dddm_sharded
from .training import distributional_training_step


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--config", type=str, default=None, help="Optional YAML config (same overlay rule as the reference)")
    p.add_argument("--data-dir", type=str, default="./data")
    p.add_argument("--out", type=str, default="./cifar10_dit_out")
    p.add_argument("--epochs", type=int, default=10)
    p.add_argument("--batch", type=int, default=128,
                   help="batch PER GPU when given on the command line; the `batch` key of a YAML config is the reference's "
                        "GLOBAL batch (train_cifar10_dit.py is single-device) and is divided by the number of ranks")
    p.add_argument("--global-batch", type=int, default=0,
                   help="global batch, divided evenly over the ranks (overrides --batch; 0 = use --batch per GPU)")
    p.add_argument("--lr", type=float, default=1e-4)
    p.add_argument("--weight-decay", type=float, default=0.01)
    p.add_argument("--beta", type=float, default=0.1)
    p.add_argument("--lam", type=float, default=1.0)
    p.add_argument("--m", type=int, default=8)
    p.add_argument("--w-bias", type=float, default=0.0, dest="w_bias")
    p.add_argument("--grad-clip", type=float, default=1.0)
    p.add_argument("--ckpt-every", type=int, default=1)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--image-size", type=int, default=32)
    p.add_argument("--patch-size", type=int, default=4)
    p.add_argument("--embed-dim", type=int, default=384)
    p.add_argument("--depth", type=int, default=8)
    p.add_argument("--heads", type=int, default=6)
    p.add_argument("--time-embed", type=int, default=256)
    p.add_argument("--mlp-ratio", type=float, default=4.0)
    p.add_argument("--sample-batch", type=int, default=64)
    p.add_argument("--sample-steps", type=int, default=20)
    p.add_argument("--eps-churn", type=float, default=1.0)
    # reference flags kept so that its command lines and configs/cifar10_dit.yaml load unchanged
    # (train_cifar10_dit.py:376-395); evaluation with Inception features, W&B and the CPU data-loader workers are outside
    # the hot path: the values are accepted, recorded in the checkpoint config, and reported as ignored at start-up
    p.add_argument("--device", type=str, default="cuda", help="must be a CUDA device (there is no CPU path)")
    p.add_argument("--workers", type=int, default=4)
    p.add_argument("--no-augment", action="store_true")
    p.add_argument("--eval-every", type=int, default=0)
    p.add_argument("--eval-batch", type=int, default=256)
    p.add_argument("--eval-samples", type=int, default=1024)
    p.add_argument("--fid-samples", type=int, default=10000)
    p.add_argument("--mmd-samples", type=int, default=2048)
    p.add_argument("--mmd-sigma", type=float, default=1.0)
    p.add_argument("--wandb", action="store_true")
    p.add_argument("--wandb-project", type=str, default="dddm")
    p.add_argument("--wandb-name", type=str, default=None)
    # launcher-only flags
    p.add_argument("--synthetic", action="store_true", help="CIFAR-shaped random images resident on the GPU (no dataset)")
    p.add_argument("--steps-per-epoch", type=int, default=100, help="with --synthetic")
    p.add_argument("--precision", choices=["fp32", "tf32", "bf16"], default="bf16",
                   help="backbone matmul precision (the loss kernels always accumulate in fp32)")
    p.add_argument("--log-every", type=int, default=20)
    p.add_argument("--cuda-graph", dest="cuda_graph", action="store_true", default=None,
                   help="capture the whole training step in one CUDA graph (default: on with --synthetic)")
    p.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false")
    p.add_argument("--grad-buckets", type=int, default=4,
                   help="data parallel: the flat gradient is all-reduced in this many buckets, each launched from a "
                        "backward hook on a side stream as soon as its last gradient exists (0/1 = one all-reduce after "
                        "the backward)")
    p.add_argument("--graph-nccl", dest="graph_nccl", action="store_true", default=None,
                   help="several GPUs: capture the NCCL all-reduces INSIDE one whole-step CUDA graph (bucketed overlap "
                        "included) instead of two graphs with eager collectives between them")
    p.add_argument("--no-graph-nccl", dest="graph_nccl", action="store_false")
    return p


def apply_yaml(parser: argparse.ArgumentParser, args: argparse.Namespace) -> None:
    """A YAML key overrides a flag only while the flag still has its parser default; unknown keys raise
    (same rule as the reference, ``train_cifar10_dit.py:67-78``)."""
    if not args.config:
        return
    import yaml

    with open(args.config, "r", encoding="utf-8") as f:
        cfg = yaml.safe_load(f) or {}
    for key, value in cfg.items():
        dest = key.replace("-", "_")
        if not hasattr(args, dest):
            raise ValueError(f"Unknown config key: {key}")
        if dest == "batch":  # the reference's batch is the global one (App. D.1: 256 = 4 x 64)
            if args.global_batch == 0 and args.batch == parser.get_default("batch"):
                args.global_batch = int(value)
            continue
        if getattr(args, dest) == parser.get_default(dest):
            setattr(args, dest, value)


IGNORED_REFERENCE_FLAGS = ("eval_every", "eval_batch", "eval_samples", "fid_samples", "mmd_samples", "mmd_sigma", "wandb",
                           "wandb_project", "wandb_name")


def resolve_batch(args: argparse.Namespace, world: int) -> list:
    """Per-GPU batch from --global-batch / YAML `batch`; returns the notices to print on rank 0."""
    notes = []
    if not str(args.device).startswith("cuda"):
        raise ValueError(f"--device {args.device}: the launcher runs on CUDA devices only (one process per GPU)")
    if args.global_batch:
        if args.global_batch % world:
            raise ValueError(f"global batch {args.global_batch} is not divisible by {world} ranks")
        args.batch = args.global_batch // world
        notes.append(f"global batch {args.global_batch} -> {args.batch} per GPU on {world} rank(s)")
    if args.eval_every or args.wandb:
        notes.append("evaluation (FID / image MMD) and W&B logging are outside this launcher: --eval-* / --fid-* / "
                     "--mmd-* / --wandb* are accepted and ignored (ddm_b200.rbf_mmd2 and sample_dddm_sharded are the "
                     "evaluation-side kernels)")
    return notes


def init_distributed():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, dev


def _flatten_(tensors, dtype) -> torch.Tensor:
    """Re-home ``tensors`` (parameters) as views of ONE contiguous buffer and return it."""
    flat = torch.empty(sum(t.numel() for t in tensors), dtype=dtype, device=tensors[0].device)
    off = 0
    for t in tensors:
        n = t.numel()
        flat[off:off + n].copy_(t.detach().reshape(-1))
        t.data = flat[off:off + n].view_as(t)
        off += n
    return flat


class Trainer:
    """Model + optimizer + one data-parallel training step (used by main() and by bench.py).

    Layout (no host synchronisation, no per-tensor launches, whole step in one CUDA graph):

    * fp32 master parameters live in ONE flat buffer (AdamW state follows); with ``--precision bf16`` the
      backbone computes on a bf16 shadow copy (also one flat buffer, refreshed by one cast kernel per step),
      activations and gradients are bf16, the loss kernels accumulate in fp32;
    * gradients are views of ONE flat buffer: data parallelism is a single NCCL all-reduce(AVG) of that buffer
      after the backward (29 MB bf16 / 58 MB fp32 over NVLink/NVSwitch: ~0.1-0.3 ms against a >15 ms step, so
      bucketed overlap buys nothing here), followed by global-norm clipping with the coefficient kept on
      the device and fused AdamW;
    * with ``--cuda-graph`` (default for --synthetic) the step is captured once and replayed: on one GPU as ONE
      graph (K4, K2c, backbone forward, K1, backward, clip, AdamW, shadow refresh); on several GPUs as two graphs
      with the two NCCL all-reduces (the w-sum float, the flat gradient) issued eagerly between them.
    """

    def __init__(self, args, dev: torch.device, world: int, module: torch.nn.Module | None = None, loss_fn=None):
        """``module`` / ``loss_fn(model, x0) -> (loss, packed_metrics)`` default to the DiT and the DDDM step; the
        CPU tests pass a small model and loss to exercise the flat-buffer / all-reduce / clip logic on gloo."""
        self.args, self.dev, self.world = args, dev, world
        torch.manual_seed(args.seed)  # identical initial weights on every rank
        if module is None:
            module = DDDMDiT(img_size=args.image_size, patch_size=args.patch_size, in_channels=6, out_channels=3,
                             embed_dim=args.embed_dim, depth=args.depth, num_heads=args.heads,
                             time_embed_dim=args.time_embed, mlp_ratio=args.mlp_ratio)
        self.module = module.to(dev)
        self.module.train()
        self.loss_fn = loss_fn or self._dddm_loss
        self.bf16 = args.precision == "bf16"
        torch.backends.cuda.matmul.allow_tf32 = args.precision != "fp32"
        torch.backends.cudnn.allow_tf32 = args.precision != "fp32"
        master = [p for p in self.module.parameters() if p.requires_grad]
        self.flat_master = _flatten_(master, torch.float32)
        if world > 1:
            dist.broadcast(self.flat_master, src=0)  # identical start even if the caller seeded the ranks differently
        self.flat_master_grad = torch.zeros_like(self.flat_master)
        if self.bf16:
            self.model = copy.deepcopy(self.module).to(torch.bfloat16)
            shadow = [p for p in self.model.parameters() if p.requires_grad]
            self.flat_shadow = _flatten_(shadow, torch.bfloat16)
            self.flat_grad = torch.zeros_like(self.flat_shadow)
            grad_owner = shadow
        else:
            self.model = self.module
            self.flat_grad = self.flat_master_grad
            grad_owner = master
        off = 0
        for p, q in zip(grad_owner, master):
            n = p.numel()
            p.grad = self.flat_grad[off:off + n].view_as(p)          # autograd accumulates in place
            q.grad = self.flat_master_grad[off:off + n].view_as(q)   # what the optimizer reads
            off += n
        # ---- gradient buckets: contiguous slices of the flat buffer, all-reduced from backward hooks (world > 1) ----
        self._buckets = []          # (start, end) in elements of flat_grad
        self._bucket_left = []      # gradients still missing per bucket in the current backward
        self._bucket_total = []
        self._hooks_on = False
        self._comm_stream = None
        nb = int(getattr(args, "grad_buckets", 4) or 0)
        if world > 1 and nb > 1:
            total = self.flat_grad.numel()
            bounds, owner_bucket, off, b = [0], [], 0, 0
            for prm in grad_owner:
                owner_bucket.append(b)
                off += prm.numel()
                if off >= (b + 1) * total / nb and b < nb - 1 and off < total:
                    bounds.append(off)
                    b += 1
            bounds.append(total)
            self._buckets = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]
            self._bucket_total = [owner_bucket.count(i) for i in range(len(self._buckets))]
            self._bucket_left = list(self._bucket_total)
            self._comm_stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None  # CPU (gloo tests): in line
            for prm, bi in zip(grad_owner, owner_bucket):
                prm.register_post_accumulate_grad_hook(self._make_bucket_hook(bi))
        want_graph = getattr(args, "cuda_graph", None)
        self.use_graph = dev.type == "cuda" and bool(getattr(args, "synthetic", False) if want_graph is None else want_graph)
        self.opt = torch.optim.AdamW(master, lr=args.lr, weight_decay=args.weight_decay, fused=dev.type == "cuda",
                                     capturable=self.use_graph)
        # Several ranks: either the step is split into two graphs around eager collectives (default: robust across
        # NCCL / driver versions), or --graph-nccl captures the collectives too (ONE graph; the bucketed all-reduces then
        # overlap the tail of the backward inside the graph)
        self.graph_nccl = bool(getattr(args, "graph_nccl", None) or os.environ.get("DDDM_GRAPH_NCCL") == "1") and world > 1
        self.split_graph = (self.use_graph and (world > 1 or os.environ.get("DDDM_SPLIT_GRAPH") == "1") and loss_fn is None
                            and not self.graph_nccl)
        self._graph = self._graph_update = None
        self._static_x0 = self._static_t = self._static_wsum = None
        self._static_metrics = None
        torch.manual_seed(args.seed + 1 + (dist.get_rank() if world > 1 else 0))  # per-rank data / noise streams

    # -- one optimisation step, all on the current stream, no host synchronisation -------------------------
    def _dddm_loss(self, model, x0, **kw):
        a = self.args
        loss, metrics = distributional_training_step(model, x0, m=a.m, beta=a.beta, lam=a.lam, w_bias=a.w_bias,
                                                     sync_metrics=False, **kw)
        return loss, metrics.tensor

    # -- the step in two device-side phases with the gradient all-reduce between them; no host synchronisation ----
    def _make_bucket_hook(self, bi: int):
        def hook(_param):
            if not self._hooks_on:
                return
            self._bucket_left[bi] -= 1
            if self._bucket_left[bi] == 0:
                self._reduce_bucket(bi)
        return hook

    def _reduce_bucket(self, bi: int) -> None:
        """All-reduce(AVG) of one bucket on the communication stream, ordered after everything the backward has queued
        so far (all of the bucket's gradients).  Under stream capture the dependency becomes a graph edge."""
        s, e = self._buckets[bi]
        if self._comm_stream is None:
            dist.all_reduce(self.flat_grad[s:e], op=dist.ReduceOp.AVG)
        else:
            main = torch.cuda.current_stream(self.dev)
            self._comm_stream.wait_stream(main)
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(self.flat_grad[s:e], op=dist.ReduceOp.AVG)
        self._bucket_left[bi] = -1  # done

    def _fwd_bwd(self, x0: torch.Tensor, overlap: bool = False, **kw) -> torch.Tensor:
        """Forward + backward.  ``overlap``: the gradient buckets are all-reduced from backward hooks on the communication
        stream while the rest of the backward runs; on return the current stream has joined it (gradients are global)."""
        self.flat_grad.zero_()
        loss, packed = self.loss_fn(self.model, x0, **kw)
        if overlap and self._buckets:
            self._bucket_left = list(self._bucket_total)
            self._hooks_on = True
            try:
                loss.backward()
            finally:
                self._hooks_on = False
            for bi, left in enumerate(self._bucket_left):  # parameters without a gradient this step: reduce what is left
                if left >= 0:
                    self._reduce_bucket(bi)
            if self._comm_stream is not None:
                torch.cuda.current_stream(self.dev).wait_stream(self._comm_stream)
        else:
            loss.backward()
            if overlap and self.world > 1:
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG)
        return packed

    def _update(self) -> None:
        a = self.args
        if self.bf16:
            self.flat_master_grad.copy_(self.flat_grad)
        if a.grad_clip is not None and a.grad_clip > 0:  # clip_grad_norm_ (train_cifar10_dit.py:167-168), device-side
            coef = (a.grad_clip / (torch.linalg.vector_norm(self.flat_master_grad) + 1e-6)).clamp(max=1.0)
            self.flat_master_grad.mul_(coef)
        self.opt.step()
        if self.bf16:
            self.flat_shadow.copy_(self.flat_master)

    def _step_impl(self, x0: torch.Tensor) -> torch.Tensor:
        packed = self._fwd_bwd(x0, overlap=self.world > 1)
        self._update()
        return packed

    # -- CUDA graphs.  One GPU: the whole step is ONE graph.  Several GPUs: NCCL stays outside the graphs (eager
    #    collectives between replays are robust across NCCL/driver versions): [rand t, K4, all-reduce w-sum] ->
    #    graph A (K2c, backbone forward, K1, backward) -> all-reduce(flat gradient) -> graph B (clip, AdamW, refresh).
    def _weights(self, batch: int) -> None:
        from . import ops

        self._static_t.copy_(torch.rand(batch, device=self.dev, dtype=self._static_t.dtype))  # first draw of the step
        _, w = ops.sigmoid_weight_sum(self._static_t, float(self.args.w_bias))
        if self.world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
        self._static_wsum.copy_(w)

    def _split_step_eager(self, x0: torch.Tensor) -> torch.Tensor:
        self._weights(x0.shape[0])
        packed = self._fwd_bwd(x0, overlap=self.world > 1, t=self._static_t, weight_sum=self._static_wsum)
        self._update()
        return packed

    def _whole_step_nccl(self, x0: torch.Tensor) -> torch.Tensor:
        """--graph-nccl: the data-parallel step with both collectives in line (capturable as ONE graph)."""
        self._weights(x0.shape[0])
        packed = self._fwd_bwd(x0, overlap=True, t=self._static_t, weight_sum=self._static_wsum)
        self._update()
        return packed

    def _snapshot(self):
        state = {k: {n: (v.clone() if torch.is_tensor(v) else v) for n, v in st.items()} for k, st in self.opt.state.items()}
        return self.flat_master.clone(), state, torch.cuda.get_rng_state(self.dev)

    def _restore(self, snap) -> None:
        """Undo the warm-up steps: they exist to build library plans, allocator pools, optimizer state tensors and NCCL
        communicators before the capture, not to train."""
        master, state, rng = snap
        self.flat_master.copy_(master)
        if self.bf16:
            self.flat_shadow.copy_(self.flat_master)
        for k, st in self.opt.state.items():
            for n, v in st.items():
                if torch.is_tensor(v):
                    old = state.get(k, {}).get(n)
                    v.copy_(old) if old is not None else v.zero_()  # no earlier state: fresh = zeros, step 0
        torch.cuda.set_rng_state(rng, self.dev)

    def _capture(self, x0: torch.Tensor) -> None:
        self._static_x0 = x0.clone()
        self._static_t = torch.zeros(x0.shape[0], device=self.dev, dtype=x0.dtype)
        self._static_wsum = torch.zeros(1, device=self.dev)
        eager = self._split_step_eager if self.split_graph else (self._whole_step_nccl if self.graph_nccl else self._step_impl)
        snap = self._snapshot()
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(3):  # warm-up: cuBLAS/cuDNN plans, optimizer state, NCCL communicators
                eager(self._static_x0)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self._restore(snap)
        self._graph = torch.cuda.CUDAGraph()
        if self.graph_nccl:
            # thread-local capture mode: the NCCL watchdog thread's event queries must not invalidate the capture
            with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                self._static_metrics = self._whole_step_nccl(self._static_x0)
            return
        if not self.split_graph:
            with torch.cuda.graph(self._graph):
                self._static_metrics = self._step_impl(self._static_x0)
            return
        with torch.cuda.graph(self._graph):
            self._static_metrics = self._fwd_bwd(self._static_x0, t=self._static_t, weight_sum=self._static_wsum)
        self._graph_update = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph_update):
            self._update()

    def step(self, x0: torch.Tensor):
        from .training import DeferredMetrics

        if not self.use_graph:
            return DeferredMetrics(self._step_impl(x0))
        if self._graph is None:
            self._capture(x0)
        self._static_x0.copy_(x0, non_blocking=True)
        if self.split_graph:
            self._weights(x0.shape[0])
            self._graph.replay()
            if self.world > 1:
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG)
            self._graph_update.replay()
        else:
            self._graph.replay()
        return DeferredMetrics(self._static_metrics.clone())

    def synthetic_batch(self) -> torch.Tensor:
        a = self.args
        return torch.rand(a.batch, 3, a.image_size, a.image_size, device=self.dev) * 2.0 - 1.0

    def save(self, path: str) -> None:
        torch.save({"model": self.module.state_dict(), "config": vars(self.args)}, path)


def measure_throughput(args, dev, world, steps: int, warmup: int) -> dict:
    """Images/s of the DP training step on synthetic CIFAR-shaped data (device-timed, max over ranks)."""
    flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        return _measure_throughput(args, dev, world, steps, warmup)
    finally:  # the Trainer sets the process-wide TF32 switches for its precision mode; do not leak them to the caller
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = flags


def _measure_throughput(args, dev, world, steps: int, warmup: int) -> dict:
    tr = Trainer(args, dev, world)
    x0 = tr.synthetic_batch()
    for _ in range(max(warmup, 1)):
        tr.step(x0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        metrics = tr.step(x0)
    e1.record()
    e1.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t)
    return {"img_per_s": world * args.batch * steps / dt, "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "global_batch": world * args.batch, "m": args.m, "precision": args.precision, "loss": metrics["loss"],
            "model": "DDDMDiT(default, 14.5M params)", "optimizer": "AdamW(fused)", "n_gpus": world}


def dp_parity(args, dev, world, batch_per_rank: int = 16) -> dict:
    """Multi-rank parity ON HARDWARE: one step of the NCCL launcher in its production form (two CUDA graphs with the
    eager all-reduces of sum_b w(t_b) and of the flat gradient between them) against the reference recipe run by ONE
    process on the GLOBAL batch (``train_cifar10_dit.py:152-169``: zero_grad -> backward -> clip_grad_norm_, with the
    loss of ``dddm/training.py:84-85``).  Every rank records the noise its step consumed (the CUDA RNG state is saved
    before the step and the draws t, eps, xi are replayed in the step's order afterwards), rank 0 gathers all shards
    once, recomputes loss and clipped gradient from the same weights through the single-GPU path and reports the
    relative differences.  fp32 backbone unless ``args.precision`` says otherwise (expect <= 1e-6 in fp32)."""
    import copy as _copy

    flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    a = _copy.copy(args)
    a.batch, a.synthetic = batch_per_rank, True
    if getattr(a, "cuda_graph", None) is None:
        a.cuda_graph = True
    try:
        tr = Trainer(a, dev, world)
        x0 = tr.synthetic_batch()
        tr.step(x0)  # captures the graphs (its warm-up steps are rolled back), then runs one step
        torch.cuda.synchronize(dev)
        w_before = tr.flat_master.clone()
        rng = torch.cuda.get_rng_state(dev)
        metrics = tr.step(x0)
        torch.cuda.synchronize(dev)
        g_dp = tr.flat_master_grad.clone()  # all-reduced (AVG) and clipped, as the optimizer saw it
        loss_dp = metrics.tensor[:1].clone().double()
        if world > 1:
            dist.all_reduce(loss_dp, op=dist.ReduceOp.AVG)  # the global loss is the mean of the rank losses (equal shards)
        # the noise this rank's step consumed, in the step's order: rand(B) (launcher), randn_like(x0), randn(B, m, ...)
        torch.cuda.set_rng_state(rng, dev)
        t = torch.rand(a.batch, device=dev, dtype=x0.dtype)
        eps = torch.randn_like(x0)
        xi = torch.randn((a.batch, a.m, *x0.shape[1:]), device=dev, dtype=x0.dtype)

        def gather(v):
            if world == 1:
                return v
            parts = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(parts, v.contiguous())
            return torch.cat(parts, dim=0)

        X0, T, E, XI = gather(x0), gather(t), gather(eps), gather(xi)
        out = {"ranks": world, "batch_per_rank": a.batch, "global_batch": a.batch * world, "precision": a.precision,
               "launcher_form": ("two CUDA graphs + eager NCCL all-reduces" if tr.split_graph else
                                 (f"ONE CUDA graph with the NCCL all-reduces captured ({len(tr._buckets) or 1} gradient buckets "
                                  "launched from backward hooks on a side stream)" if tr.graph_nccl else
                                  ("one CUDA graph (1 GPU)" if tr.use_graph else "eager")))}
        if (dist.get_rank() if world > 1 else 0) == 0:
            ref = _copy.deepcopy(tr.module)
            flat_ref = _flatten_([p for p in ref.parameters() if p.requires_grad], torch.float32)
            flat_ref.copy_(w_before)
            if tr.bf16:
                ref = ref.to(torch.bfloat16)
            ref.train()
            loss, _ = distributional_training_step(ref, X0, m=a.m, beta=a.beta, lam=a.lam, w_bias=a.w_bias, t=T, eps=E,
                                                   xi=XI, global_weight=False, sync_metrics=False)
            grads = torch.autograd.grad(loss, [p for p in ref.parameters() if p.requires_grad])
            g_ref = torch.cat([g.reshape(-1).float() for g in grads])
            if a.grad_clip is not None and a.grad_clip > 0:
                g_ref = g_ref * (a.grad_clip / (torch.linalg.vector_norm(g_ref) + 1e-6)).clamp(max=1.0)
            loss = loss.detach()
            out.update({
                "loss_dp": float(loss_dp), "loss_global": float(loss),
                "loss_rel": abs(float(loss_dp) - float(loss)) / max(abs(float(loss)), 1e-30),
                "grad_max_rel": float((g_dp - g_ref).abs().max() / g_ref.abs().max().clamp_min(1e-30)),
                "grad_l2_rel": float(torch.linalg.vector_norm(g_dp - g_ref) / torch.linalg.vector_norm(g_ref).clamp_min(1e-30)),
                "grad_norm_after_clip": float(torch.linalg.vector_norm(g_ref)),
            })
        if world > 1:
            dist.barrier()
        return out
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = flags


def build_train_loader(args, world: int, rank: int, dataset=None):
    """The real-CIFAR input pipeline of the reference (`dddm/data.py:195-247`, called from `train_cifar10_dit.py:104-113`):
    reflect-padded random crop of 32 + horizontal flip unless `--no-augment`, resize when `--image-size` is not 32, `ToTensor`,
    normalisation to [-1, 1]; shuffled, `drop_last=True`, `--workers` workers, pinned memory.  Differences, both forced by one
    process per GPU: every rank reads its own `DistributedSampler` shard (re-seeded per epoch by the caller) and `args.batch` is
    the PER-GPU batch (`resolve_batch`).  The dataset is never downloaded (no network on the training boxes): `--data-dir` must
    hold `cifar-10-batches-py`.  `dataset` replaces `datasets.CIFAR10` (any dataset of (PIL image, label) pairs: the CPU
    test)."""
    try:
        from torchvision import datasets, transforms
    except ImportError as exc:
        raise RuntimeError("torchvision is needed for CIFAR-10; pass --synthetic to train on random images") from exc
    tfms = []
    if not args.no_augment:
        tfms += [transforms.RandomCrop(32, padding=4, padding_mode="reflect"), transforms.RandomHorizontalFlip()]
    if args.image_size != 32:
        tfms.append(transforms.Resize(args.image_size))
    tfms += [transforms.ToTensor(), transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))]
    tf = transforms.Compose(tfms)
    if dataset is None:
        ds = datasets.CIFAR10(args.data_dir, train=True, download=False, transform=tf)
    else:
        ds = _Transformed(dataset, tf)
    sampler = (torch.utils.data.distributed.DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=True, seed=args.seed)
               if world > 1 else None)
    return torch.utils.data.DataLoader(ds, batch_size=args.batch, sampler=sampler, shuffle=sampler is None,
                                       num_workers=args.workers, pin_memory=torch.cuda.is_available(), drop_last=True)


class _Transformed(torch.utils.data.Dataset):
    """(image, label) dataset with the training transform applied on access, like `datasets.CIFAR10(transform=...)`."""

    def __init__(self, base, transform):
        self.base, self.transform = base, transform

    def __len__(self):
        return len(self.base)

    def __getitem__(self, i):
        img, label = self.base[i]
        return self.transform(img), label


def main(argv=None) -> None:
    parser = build_parser()
    args = parser.parse_args(argv)
    apply_yaml(parser, args)
    if args.m < 2:
        parser.error("m must be >= 2 for the generalized energy score")
    world, rank, dev = init_distributed()
    for note in resolve_batch(args, world):
        if rank == 0:
            print(f"[ddm_b200.launcher] {note}", flush=True)
    os.makedirs(args.out, exist_ok=True)
    tr = Trainer(args, dev, world)

    loader = None if args.synthetic else build_train_loader(args, world, rank)

    history, gstep = [], 0
    for epoch in range(1, args.epochs + 1):
        tr.model.train()
        if loader is not None and world > 1:
            loader.sampler.set_epoch(epoch)
        batches = ((x.to(dev, non_blocking=True) for x, _ in loader) if loader is not None else
                   (tr.synthetic_batch() for _ in range(args.steps_per_epoch)))
        t0, seen, pending = time.perf_counter(), 0, []
        for x0 in batches:
            pending.append(tr.step(x0).tensor)
            gstep += 1
            seen += x0.shape[0] * world
            if gstep % args.log_every == 0:
                packed = torch.stack(pending).mean(dim=0)  # loss, conf, inter, weight — still on the device
                if world > 1:
                    dist.all_reduce(packed, op=dist.ReduceOp.AVG)
                vals = packed.tolist()  # the only host synchronisation of the interval
                pending.clear()
                if rank == 0:
                    rec = dict(step=gstep, epoch=epoch, loss=vals[0], confidence=vals[1], interaction=vals[2],
                               weight=vals[3], img_per_s=seen / (time.perf_counter() - t0))
                    history.append(rec)
                    print(json.dumps(rec), flush=True)
        if rank == 0 and args.ckpt_every > 0 and epoch % args.ckpt_every == 0:
            tr.save(os.path.join(args.out, f"ckpt_epoch_{epoch:04d}.pt"))
    if rank == 0:
        tr.save(os.path.join(args.out, "model_final.pt"))
        with open(os.path.join(args.out, "train_history.json"), "w", encoding="utf-8") as f:
            json.dump(history, f, indent=2)
    if args.sample_batch > 0:
        n = -(-args.sample_batch // world) * world
        x = sample_dddm_sharded(tr.module, n, steps=args.sample_steps, eps_churn=args.eps_churn,
                                data_shape=(3, args.image_size, args.image_size), seed=args.seed)
        if rank == 0:
            torch.save(x[: args.sample_batch].cpu(), os.path.join(args.out, "samples.pt"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
