#!/usr/bin/env python
"""Small shapes through every kernel variant, for `compute-sanitizer --tool memcheck` (one tool per gpurun call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ddm_b200
from ddm_b200 import _cabi, ops

dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(0)


def energy(B, m, D, dtype, variant=0, cluster=0):
    _cabi.set_tuning("energy.variant", variant)
    _cabi.set_tuning("energy.cluster", cluster)
    x0 = torch.randn(B, D, generator=gen).clamp(-1, 1).to(dev).to(dtype)
    xh = (x0[:, None].float().cpu() + 0.05 * torch.randn(B, m, D, generator=gen)).to(dev).to(dtype)
    w = torch.full((1,), 0.5 * B, device=dev)
    try:
        out, g = ops.energy_fused(xh, x0, w, 1.0 / B, 0.1, 1.0, True)
        a = xh.clone().requires_grad_(True)
        c = x0.clone().requires_grad_(True)
        conf, inter = ddm_b200.generalized_energy_terms(a, c, 0.1, 1.0)
        (conf - 0.07 * inter).backward()
        return _cabi.describe_energy(B, m, D, "f32" if dtype == torch.float32 else "bf16"), float(out[0])
    except _cabi.DDDMError as e:
        return f"unsupported plan ({e.status})", None


for dtype in (torch.float32, torch.bfloat16):
    for (B, m, D, variant, cluster) in ((5, 8, 3072, 0, 0), (3, 8, 3072, 3, 2), (4, 4, 12288, 0, 0), (3, 8, 259, 0, 0),
                                        (2, 8, 2, 0, 0), (3, 16, 3072, 0, 0), (2, 32, 3072, 0, 0), (2, 32, 3072, 4, 8),
                                        (2, 24, 1024, 0, 0), (2, 12, 100, 0, 0), (2, 8, 3072, 2, 0), (2, 8, 3072, 1, 4)):
        print(dtype, B, m, D, energy(B, m, D, dtype, variant, cluster), flush=True)
_cabi.set_tuning("energy.variant", 0)
_cabi.set_tuning("energy.cluster", 0)
x0 = (torch.rand(4, 3, 32, 32, generator=gen) * 2 - 1).to(dev)
t, eps, xi = torch.rand(4, generator=gen).to(dev), torch.randn(4, 3, 32, 32, generator=gen).to(dev), torch.randn(4, 3, 3, 32, 32, generator=gen).to(dev)
ops.forward_marginal_concat(x0, t, eps, xi, True, 4)
ops.forward_marginal_expand(x0, t, eps, 3, True)
ops.bridge_step(x0, eps, eps, t[:1], t[1:2], 0.7)
ops.sigmoid_weight_sum(t, 0.1)
a = torch.randn(70, 384, device=dev, dtype=torch.bfloat16, requires_grad=True)
wt, bs = torch.ones(384, device=dev, dtype=torch.bfloat16, requires_grad=True), torch.zeros(384, device=dev, dtype=torch.bfloat16, requires_grad=True)
ops.layer_norm(a, wt, bs, 1e-5)[0].sum().backward()
ops.colsum(a.detach())
print("mmd", float(ddm_b200.rbf_mmd2(torch.randn(300, 5, device=dev), torch.randn(257, 5, device=dev), 1.3)))
torch.cuda.synchronize()
print("sanitize_small done")
