"""Data-parallel DDDM training on one box of B200s: one process per GPU, NCCL over NVLink.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m ddm_b200.launcher --synthetic ...

The reference's ``train_cifar10_dit.py`` is single-process, single-device (``:99-100``); this is the
data-parallel launcher BASELINE.json asks for, with the same flag names and defaults
(``train_cifar10_dit.py:362-398``), the same per-step order (``:152-169``: step -> zero_grad ->
backward -> clip_grad_norm_ -> AdamW.step) and the same checkpoint payload (``:32-37``).

Per step and per rank:
  K4  w-sum of the local t  ->  1-float NCCL all-reduce (async, overlaps the backbone forward)
  K2  x_t = alpha x0 + sigma eps written m-fold straight into the backbone input
  backbone forward (PyTorch DDP; bf16 autocast optional)
  K1  fused energy-score loss forward + backward with the GLOBAL weight
  backbone backward with DDP's bucketed gradient all-reduce (58 MB fp32) overlapped
  fused AdamW; metrics stay on the device and are read back packed, once per log interval.
The batch is sharded by row (``--batch`` is per GPU, BASELINE config 4: 128/GPU); there is no
other parallelism axis (14.5 M parameters, 64-token sequences).
"""
from __future__ import annotations

import argparse
import contextlib
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
    p.add_argument("--batch", type=int, default=128, help="per-GPU batch")
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
    # launcher-only flags
    p.add_argument("--synthetic", action="store_true", help="CIFAR-shaped random images resident on the GPU (no dataset)")
    p.add_argument("--steps-per-epoch", type=int, default=100, help="with --synthetic")
    p.add_argument("--precision", choices=["fp32", "tf32", "bf16"], default="bf16",
                   help="backbone matmul precision (the loss kernels always accumulate in fp32)")
    p.add_argument("--log-every", type=int, default=20)
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
        if getattr(args, dest) == parser.get_default(dest):
            setattr(args, dest, value)


def init_distributed():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, dev


def set_precision(precision: str):
    torch.backends.cuda.matmul.allow_tf32 = precision != "fp32"
    torch.backends.cudnn.allow_tf32 = precision != "fp32"
    if precision == "bf16":
        return lambda: torch.autocast("cuda", dtype=torch.bfloat16)
    return contextlib.nullcontext


class Trainer:
    """Model + optimizer + one data-parallel training step (used by main() and by bench.py)."""

    def __init__(self, args, dev: torch.device, world: int):
        self.args, self.dev, self.world = args, dev, world
        torch.manual_seed(args.seed)  # identical initial weights on every rank
        model = DDDMDiT(img_size=args.image_size, patch_size=args.patch_size, in_channels=6, out_channels=3,
                        embed_dim=args.embed_dim, depth=args.depth, num_heads=args.heads,
                        time_embed_dim=args.time_embed, mlp_ratio=args.mlp_ratio).to(dev)
        self.module = model
        if world > 1:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], gradient_as_bucket_view=True)
        self.model = model
        self.opt = torch.optim.AdamW(self.module.parameters(), lr=args.lr, weight_decay=args.weight_decay, fused=True)
        self.autocast = set_precision(args.precision)
        torch.manual_seed(args.seed + 1 + (dist.get_rank() if world > 1 else 0))  # per-rank data / noise streams

    def step(self, x0: torch.Tensor):
        a = self.args
        with self.autocast():
            loss, metrics = distributional_training_step(self.model, x0, m=a.m, beta=a.beta, lam=a.lam,
                                                         w_bias=a.w_bias, sync_metrics=False)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if a.grad_clip is not None and a.grad_clip > 0:
            torch.nn.utils.clip_grad_norm_(self.module.parameters(), a.grad_clip)
        self.opt.step()
        return metrics

    def synthetic_batch(self) -> torch.Tensor:
        a = self.args
        return torch.rand(a.batch, 3, a.image_size, a.image_size, device=self.dev) * 2.0 - 1.0

    def save(self, path: str) -> None:
        torch.save({"model": self.module.state_dict(), "config": vars(self.args)}, path)


def measure_throughput(args, dev, world, steps: int, warmup: int) -> dict:
    """Images/s of the DP training step on synthetic CIFAR-shaped data (device-timed, max over ranks)."""
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


def main(argv=None) -> None:
    parser = build_parser()
    args = parser.parse_args(argv)
    apply_yaml(parser, args)
    if args.m < 2:
        parser.error("m must be >= 2 for the generalized energy score")
    world, rank, dev = init_distributed()
    os.makedirs(args.out, exist_ok=True)
    tr = Trainer(args, dev, world)

    loader = None
    if not args.synthetic:
        try:
            from torchvision import datasets, transforms
        except ImportError as exc:
            raise RuntimeError("torchvision is needed for CIFAR-10; pass --synthetic to train on random images") from exc
        tf = transforms.Compose([transforms.RandomCrop(args.image_size, padding=4), transforms.RandomHorizontalFlip(),
                                 transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
        ds = datasets.CIFAR10(args.data_dir, train=True, download=False, transform=tf)
        sampler = torch.utils.data.distributed.DistributedSampler(ds) if world > 1 else None
        loader = torch.utils.data.DataLoader(ds, batch_size=args.batch, sampler=sampler, shuffle=sampler is None,
                                             num_workers=4, pin_memory=True, drop_last=True)

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
