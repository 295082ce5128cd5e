#!/usr/bin/env python
"""BASELINE config 1 end to end on the GPU kernels: the reference's toy flow (run_example.py:60-110, train_dddm at
dddm/training.py:95-170) — DDDMMLP on the 2-D bimodal GMM, batch 512, m=8, beta=0.1, lam=1, Adam lr 2e-3, then
sample_dddm with 20 steps and rbf_mmd2(sigma=1) against fresh GMM samples — with every hot-path function coming from
ddm_b200 (K4, K2, K1 register-resident variant for D=2, K3, K5).  Prints one JSON line.

    python tools/toy_gmm_e2e.py [--steps 1500] [--seed 42]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ddm_b200
from ddm_b200.backbones import DDDMMLP


def sample_gmm(batch, device, sigma=0.5):
    """dddm/data.py:35-47: means (3, 3) and (-3, 3), sigma 0.5, equal weights."""
    mu = torch.tensor([[3.0, 3.0], [-3.0, 3.0]], device=device)
    pick = mu[torch.bernoulli(0.5 * torch.ones(batch, device=device)).long()]
    return pick + sigma * torch.randn(batch, 2, device=device)


def run(steps: int, seed: int, dev: str = "cuda:0", log_every: int = 250):
    torch.manual_seed(seed)
    model = DDDMMLP().to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    xref = sample_gmm(4096, dev)
    mmd0 = float(ddm_b200.rbf_mmd2(ddm_b200.sample_dddm(model, 4096, steps=20, device=dev), xref, 1.0))
    hist = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(1, steps + 1):
        x0 = sample_gmm(512, dev)
        loss, metrics = ddm_b200.distributional_training_step(model, x0, m=8, beta=0.1, lam=1.0, w_bias=0.0,
                                                              sync_metrics=False)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if step % log_every == 0 or step == steps:
            hist.append({"step": step, **{k: round(v, 5) for k, v in dict(metrics).items()}})
    torch.cuda.synchronize()
    train_s = time.perf_counter() - t0
    xgen = ddm_b200.sample_dddm(model, n_samples=4096, steps=20, device=dev)
    mmd = float(ddm_b200.rbf_mmd2(xgen, xref, 1.0))
    left = float((xgen[:, 0] < 0).float().mean())
    near = float(((xgen - torch.tensor([3.0, 3.0], device=dev)).norm(dim=1).minimum(
        (xgen - torch.tensor([-3.0, 3.0], device=dev)).norm(dim=1)) < 1.5).float().mean())
    return {"steps": steps, "seed": seed, "mmd2_untrained": mmd0, "mmd2_rbf_sigma1": mmd, "fraction_left_mode": left,
            "fraction_within_3sigma_of_a_mode": near, "train_seconds": train_s, "steps_per_s": steps / train_s,
            "history": hist}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1500)
    ap.add_argument("--seed", type=int, default=42)
    a = ap.parse_args()
    print(json.dumps(run(a.steps, a.seed)))
