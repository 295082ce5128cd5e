"""Drop-in for ``dddm/losses.py`` on the CUDA kernels (same names, arguments and return types)."""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops


def generalized_energy_terms(x0hats: torch.Tensor, x0: torch.Tensor, beta: float, lam: float
                             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Generalized energy score terms for one minibatch — reference ``dddm/losses.py:5-25``.

    ``x0hats`` [B, m, D], ``x0`` [B, D] -> (conf, inter): 0-dim tensors in the input dtype,
    differentiable w.r.t. ``x0hats`` and ``x0``.  ``lam`` is accepted and ignored exactly like the
    reference (``losses.py:6``: never read in the body).  Split forward / backward kernels (K1b):
    the forward saves the m + m(m-1)/2 squared distances per row, the backward reuses them.
    """
    del lam
    if x0hats.dim() != 3:
        raise ValueError(f"x0hats must be [B, m, D], got {tuple(x0hats.shape)}")
    if x0hats.shape[1] < 2:
        raise ValueError("m must be >= 2 to form interaction pairs")
    out, _dist = ops.energy_terms_fwd(x0hats, x0, float(beta))
    return out[0].to(x0hats.dtype), out[1].to(x0hats.dtype)


def sigmoid_weight(t: torch.Tensor, bias: float = 0.0) -> torch.Tensor:
    """w(t) = sigmoid(log(alpha^2 / (sigma^2 + 1e-12) + 1e-12) - bias) — reference ``dddm/losses.py:28-35``."""
    w, _ = ops.sigmoid_weight_sum(t, float(bias))
    return w.reshape(t.shape).to(t.dtype)
