#!/usr/bin/env python
"""Data-parallel launcher forms under torchrun: parity against the single-process global-batch recipe and DiT step time.

    torchrun --nproc-per-node N tools/dp_overlap.py --mode {split,graph_nccl,eager} [--steps 60] [--buckets 4]

split       two CUDA graphs, eager NCCL all-reduces between them (one all-reduce of the whole flat gradient)
graph_nccl  ONE CUDA graph with the collectives captured; the gradient goes in buckets launched from backward hooks on a
            side stream (they overlap the tail of the backward)
eager       no graphs; bucketed hook-launched all-reduces
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from ddm_b200 import launcher


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="split")
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--buckets", type=int, default=4)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    argv = ["--synthetic", "--precision", "fp32", "--grad-buckets", str(a.buckets)]
    argv += {"split": ["--cuda-graph", "--no-graph-nccl"], "graph_nccl": ["--cuda-graph", "--graph-nccl"],
             "eager": ["--no-cuda-graph"]}[a.mode]
    args = launcher.build_parser().parse_args(argv)
    res = {"mode": a.mode, "ranks": world, "buckets": a.buckets}
    par = launcher.dp_parity(args, dev, world)
    if rank == 0:
        res["dp_parity_fp32"] = {k: par.get(k) for k in ("loss_rel", "grad_max_rel", "grad_l2_rel", "launcher_form")}
    args = launcher.build_parser().parse_args(argv[:2] + [a.precision] + argv[3:])
    thr = launcher.measure_throughput(args, dev, world, steps=a.steps, warmup=10)
    if rank == 0:
        res["dit_" + a.precision] = {k: thr[k] for k in ("img_per_s", "ms_per_step", "steps")}
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
