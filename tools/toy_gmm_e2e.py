#!/usr/bin/env python
"""BASELINE config 1 end to end on the GPU kernels: the reference's toy flow (run_example.py:60-110, train_dddm at
dddm/training.py:95-170) — DDDMMLP on the 2-D bimodal GMM, batch 512, m=8, beta=0.1, lam=1, Adam lr 2e-3, then
sample_dddm with 20 steps and rbf_mmd2(sigma=1) against fresh GMM samples — with every hot-path function coming from
ddm_b200 (K4, K2, K1 register-resident variant for D=2, K3, K5).  Prints one JSON line.

    python tools/toy_gmm_e2e.py [--steps 1500] [--seed 42] [--reference /path/to/edluyuan_ddm]

With --reference (a checkout of the reference; absent on the GPU boxes of this project, so the default is the loop
above) nothing is re-written: the reference's OWN ``train_dddm`` (dddm/training.py:95-170), ``sample_dddm`` and
``rbf_mmd2`` run as ``run_example.py:88-101`` calls them, after ``ddm_b200.patch_reference(dddm)`` has rebound their
hot-path functions onto the kernels.
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


def run_patched_reference(ref_root: str, steps: int, seed: int, dev: str = "cuda:0"):
    """run_example.py:88-101 with the reference's own functions, patched onto ddm_b200."""
    import tempfile
    import types

    for name in ("matplotlib", "matplotlib.pyplot"):  # dddm/data.py:9 imports pyplot at module level
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, ref_root)
    import dddm

    saved = ddm_b200.patch_reference(dddm)
    try:
        cfg = dddm.TrainConfig(epochs=steps, device=dev, seed=seed)
        t0 = time.perf_counter()
        with tempfile.TemporaryDirectory() as out:
            model, history = dddm.train_dddm(cfg, outdir=out, return_history=True)
        torch.cuda.synchronize()
        train_s = time.perf_counter() - t0
        xgen = dddm.sample_dddm(model, n_samples=4096, steps=20, device=dev)
        xref = dddm.sample_gmm(4096, device=dev)
        mmd = float(dddm.rbf_mmd2(xgen, xref, sigma=1.0))
    finally:
        ddm_b200.unpatch_reference(saved)
    return {"driver": "reference train_dddm / sample_dddm / rbf_mmd2, patched", "steps": steps, "seed": seed,
            "mmd2_rbf_sigma1": mmd, "train_seconds": train_s, "steps_per_s": steps / train_s,
            "final": {k: v[-1] for k, v in history.items()}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1500)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--reference", default="", help="reference checkout: drive ITS train_dddm through patch_reference")
    a = ap.parse_args()
    if a.reference and os.path.isdir(os.path.join(a.reference, "dddm")):
        print(json.dumps(run_patched_reference(a.reference, a.steps, a.seed)))
    else:
        print(json.dumps(run(a.steps, a.seed)))
