"""CPU PyTorch port of the DDDM hot path — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's CPU path *is* eager PyTorch + autograd, and ``/root/reference`` is not present on
the GPU box, so the "reference CPU implementation" that ``bench.py`` times beside the kernels
(``cpu_baseline``, ``--impl reference``; ``kind: "port"``) is this port: the same eager tensor
program (broadcast differences, squares, reductions, boolean off-diagonal selection, ``pow``,
``mean``, autograd backward), written against the arithmetic of

* ``dddm/losses.py:5-35``   (energy terms, logistic weight)
* ``dddm/schedules.py:5-78`` (schedule, forward marginal, Gaussian bridge)
* ``dddm/training.py:64-85`` (loss combination)           — with the noise passed in
* ``dddm/sampling.py:23-31`` (Algorithm 2 loop)            — with the noise passed in

``tests/test_oracle_golden.py`` pins it (and the C oracle) to outputs of the reference's own
functions stored in ``tests/golden/``.  Never imported by ``ddm_b200``.
"""
from __future__ import annotations

import torch

POW_EPS = 1e-12  # losses.py:14,24
W_EPS = 1e-12  # losses.py:33-34
BRIDGE_EPS = 1e-8  # schedules.py:47


def _beta_power(sq: torch.Tensor, beta: float) -> torch.Tensor:
    # losses.py:11-14 / 21-24: no epsilon and no pow when beta is exactly 2.0
    return sq if beta == 2.0 else (sq + POW_EPS).pow(beta / 2.0)


def energy_terms(xhat: torch.Tensor, x0: torch.Tensor, beta: float):
    """(conf, inter) — losses.py:5-25.  xhat [B,m,D], x0 [B,D]."""
    n_draws = xhat.shape[1]
    to_target = (x0.unsqueeze(1) - xhat).pow(2).sum(dim=-1)  # [B,m]
    conf = _beta_power(to_target, beta).mean()
    between = (xhat.unsqueeze(2) - xhat.unsqueeze(1)).pow(2).sum(dim=-1)  # [B,m,m], materialises [B,m,m,D]
    off_diag = ~torch.eye(n_draws, dtype=torch.bool, device=xhat.device)
    between = between[off_diag.expand_as(between)].view(xhat.shape[0], n_draws, n_draws - 1)
    inter = _beta_power(between, beta).mean()
    return conf, inter


def sigmoid_weight(t: torch.Tensor, bias: float = 0.0) -> torch.Tensor:
    """w(t) — losses.py:28-35 with alpha = 1 - t, sigma = t (schedules.py:5-14)."""
    snr = (1.0 - t) * (1.0 - t) / (t * t + W_EPS)
    return torch.sigmoid(torch.log(snr + W_EPS) - bias)


def _pad_right(v: torch.Tensor, ndim: int) -> torch.Tensor:
    return v.reshape(v.shape + (1,) * (ndim - v.ndim)) if v.ndim < ndim else v


def forward_marginal(x0: torch.Tensor, t: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """x_t = (1 - t) x0 + t eps — schedules.py:17-25."""
    tt = _pad_right(t, x0.ndim)
    return (1.0 - tt) * x0 + tt * _pad_right(eps, x0.ndim)


def bridge(s: torch.Tensor, t: torch.Tensor, x0hat: torch.Tensor, xt: torch.Tensor, eps_churn: float = 1.0):
    """(mu, std) — schedules.py:28-78."""
    rho = s / (t + BRIDGE_EPS)
    ar = (1.0 - t) / ((1.0 - s) + BRIDGE_EPS)
    e2 = eps_churn**2
    nd = x0hat.ndim
    c_xt = e2 * _pad_right(ar * rho**2, nd) + (1.0 - e2) * _pad_right(rho, nd)
    c_x0 = _pad_right(1.0 - s, nd) * (1.0 - e2 * _pad_right(ar * rho**2, nd) - (1.0 - e2) * _pad_right(ar * rho, nd))
    inner = e2 * (ar * rho) + (1.0 - e2)
    std = ((s**2) * (1.0 - inner**2).clamp(min=0.0)).clamp(min=0.0).sqrt()
    return c_xt * xt + c_x0 * x0hat, _pad_right(std, nd)


def loss_from_draws(xhat: torch.Tensor, x0: torch.Tensor, t: torch.Tensor, *, beta: float, lam: float,
                    w_bias: float, weight: torch.Tensor | None = None):
    """training.py:77-85 given the denoiser output: (loss, conf, inter, weight)."""
    B, m = xhat.shape[0], xhat.shape[1]
    conf, inter = energy_terms(xhat.reshape(B, m, -1), x0.reshape(B, -1), beta)
    if weight is None:
        weight = sigmoid_weight(t, w_bias).mean()
    loss = weight * (conf - (lam / (2.0 * (m - 1))) * inter)
    return loss, conf, inter, weight


def training_step(model, x0: torch.Tensor, t: torch.Tensor, eps: torch.Tensor, xi: torch.Tensor, *, m: int,
                  beta: float, lam: float, w_bias: float):
    """training.py:57-85 with (t, eps, xi) passed in instead of drawn.  Returns (loss, conf, inter, weight, x0hat)."""
    if m < 2:
        raise ValueError("m must be >= 2 to form interaction pairs")
    B = x0.shape[0]
    xt = forward_marginal(x0, t, eps)
    xt_rep = xt.unsqueeze(1).expand(B, m, *xt.shape[1:]).reshape(B * m, *xt.shape[1:])
    x0hat = model(xt_rep, t.repeat_interleave(m), xi.reshape(B * m, *xt.shape[1:]))
    x0hat = x0hat.view(B, m, *x0.shape[1:])
    loss, conf, inter, weight = loss_from_draws(x0hat, x0, t, beta=beta, lam=lam, w_bias=w_bias)
    return loss, conf, inter, weight, x0hat


@torch.no_grad()
def sample(model, x_init: torch.Tensor, xis: torch.Tensor, zs: torch.Tensor, steps: int, eps_churn: float = 1.0):
    """sampling.py:20-32 with x_T = x_init, xi_k = xis[k], z_k = zs[k] passed in (index k = loop variable)."""
    grid = torch.linspace(0.0, 1.0, steps + 1, device=x_init.device)
    x = x_init
    n = x.shape[0]
    for k in reversed(range(steps)):
        s, t = grid[k], grid[k + 1]
        x0hat = model(x, t.repeat(n), xis[k])
        mu, std = bridge(s, t, x0hat, x, eps_churn)
        x = mu + std * zs[k]
    return x


def energy_fwd_bwd(xhat: torch.Tensor, x0: torch.Tensor, weight: torch.Tensor, beta: float, lam: float):
    """One timed unit of the CPU baseline: loss forward + autograd backward w.r.t. xhat."""
    xh = xhat.detach().requires_grad_(True)
    m = xh.shape[1]
    conf, inter = energy_terms(xh, x0, beta)
    loss = weight * (conf - (lam / (2.0 * (m - 1))) * inter)
    loss.backward()
    return loss.detach(), xh.grad
