"""Drop-in for ``dddm/schedules.py`` on the CUDA kernels."""
from __future__ import annotations

import torch

from . import ops


def alpha_sigma(t: torch.Tensor):
    """Flow-matching schedule alpha = 1 - t, sigma = t — reference ``dddm/schedules.py:5-14`` (trivial, stays in torch)."""
    return 1.0 - t, t


def forward_marginal_sample(x0: torch.Tensor, t: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """x_t = alpha_t x_0 + sigma_t eps — reference ``dddm/schedules.py:17-25``.

    ``t`` has one entry per leading row of ``x0``; a lower-rank ``eps`` is right-padded with
    singleton dims and broadcast, as the reference does (``schedules.py:20-21``).
    """
    while eps.ndim < x0.ndim:
        eps = eps.unsqueeze(-1)
    if eps.shape != x0.shape:
        eps = eps.expand_as(x0)
    if t.ndim == 0:
        t = t.expand(x0.shape[0])
    xt, _ = ops.forward_marginal_expand(x0, t, eps.to(x0.dtype), 0, True)
    return xt


def gaussian_bridge_mu_sigma(s: torch.Tensor, t: torch.Tensor, x0: torch.Tensor, xt: torch.Tensor,
                             eps_churn: float = 1.0):
    """Bridge transition parameters (mu, std) — reference ``dddm/schedules.py:28-78``.

    ``s``, ``t``: 0-dim or [B] tensors; returns mu like ``x0`` and std right-padded to ``x0.ndim``
    (shape [1,...,1] for scalar times, [B,1,...,1] for per-sample times), like the reference.
    """
    s = torch.as_tensor(s, device=xt.device)
    t = torch.as_tensor(t, device=xt.device)
    mu, std = ops.bridge_mu_sigma(s, t, x0, xt, float(eps_churn))
    std = std.to(x0.dtype).reshape(std.numel(), *([1] * (x0.ndim - 1)))
    return mu, std
