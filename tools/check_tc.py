#!/usr/bin/env python
"""Diagnose the tensor-core energy kernel (variant 7: m = 16 / 32 bf16 draws) against the fp64 oracle: saved squared
distances (pass 1: Gram + flags + direct fallback), loss terms, gradient (pass 2: coefficient mixing + epilogue +
post-pass), in the late / early / duplicate regimes.  Prints errors instead of asserting.

    python tools/check_tc.py [--m 32] [--D 3072] [--B 8]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from ddm_b200 import _cabi, ops


def make(B, m, D, regime, seed):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, D, generator=g).clamp(-1, 1)
    if regime == "early":
        xh = torch.randn(B, m, D, generator=g)
    elif regime == "late":
        xh = x0[:, None, :] + 0.05 * torch.randn(B, m, D, generator=g)
    elif regime == "dups":  # identical draws (zero-initialised output layer), some exact duplicates of x0
        xh = x0[:, None, :].repeat(1, m, 1) * 0.5
        xh[:, 1] = x0
        xh[:, 3] += 1e-3 * torch.randn(B, D, generator=g)
    elif regime == "mixed":  # a few near-duplicate pairs among spread draws
        xh = x0[:, None, :] + 0.3 * torch.randn(B, m, D, generator=g)
        xh[:, 5] = xh[:, 2] + 2e-3 * torch.randn(B, D, generator=g)
        xh[:, 7] = xh[:, 2]
        xh[:, 9] = x0 + 1e-3 * torch.randn(B, D, generator=g)
    return xh, x0


def ref_dist(xh, x0):
    B, m, D = xh.shape
    conf = ((xh - x0[:, None]) ** 2).sum(-1)
    iu = np.triu_indices(m, 1)
    diff = xh[:, :, None, :] - xh[:, None, :, :]
    pair = (diff ** 2).sum(-1)[:, iu[0], iu[1]]
    return np.concatenate([conf, pair], axis=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=32)
    ap.add_argument("--D", type=int, default=3072)
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--variant", type=int, default=7)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    _cabi.set_tuning("energy.variant", a.variant)
    print("kernel:", _cabi.describe_energy(a.B, a.m, a.D, "bf16"), flush=True)
    for x0_dtype in (torch.bfloat16, torch.float32):
        for regime in ("late", "early", "mixed", "dups"):
            for beta in (0.1, 1.0, 2.0):
                xh, x0 = make(a.B, a.m, a.D, regime, seed=7)
                xh = xh.to(dev).to(torch.bfloat16)
                x0 = x0.to(dev).to(x0_dtype)
                xh64, x064 = xh.double().cpu().numpy(), x0.double().cpu().numpy()
                loss, conf, inter, grad = oracle.energy_loss(xh64, x064, beta, 1.3, 0.7)
                wt = torch.tensor([0.7], device=dev)
                out, g = ops.energy_fused(xh, x0, wt, 1.0, beta, 1.3, True)
                torch.cuda.synchronize()
                out = out.cpu().numpy().astype(np.float64)
                g = g.float().cpu().numpy().astype(np.float64)
                scale = max(abs(conf), abs(inter), 1e-30)
                gerr = np.max(np.abs(g - grad)) / max(np.max(np.abs(grad)), 1e-30)
                line = (f"x0={str(x0_dtype)[6:]:8s} {regime:6s} beta={beta}: conf_err={abs(out[1] - conf) / scale:.2e} "
                        f"inter_err={abs(out[2] - inter) / scale:.2e} loss_err={abs(out[0] - loss) / (0.7 * scale):.2e} grad_err={gerr:.2e}")
                if x0_dtype == torch.bfloat16:
                    o2, dist = ops.energy_terms_fwd(xh, x0, beta)
                    d_ref = ref_dist(xh64, x064)
                    d = dist.cpu().numpy().astype(np.float64)
                    rel = np.abs(d - d_ref) / np.maximum(d_ref, 1e-30)
                    rel[d_ref == 0] = np.abs(d[d_ref == 0])
                    line += f" dist_max_rel={rel.max():.2e} (conf {rel[:, :a.m].max():.2e})"
                print(line, flush=True)
                if not np.isfinite(gerr) or gerr > 1e-2:
                    i = np.unravel_index(np.argmax(np.abs(g - grad)), g.shape)
                    print("   worst grad entry", i, g[i], grad[i], " row0 draw0 first 4:", g[0, 0, :4], grad[0, 0, :4], flush=True)
    _cabi.set_tuning("energy.variant", 0)


if __name__ == "__main__":
    main()
