"""Drop-in for ``dddm/training.py::distributional_training_step`` on the CUDA kernels."""
from __future__ import annotations

from collections.abc import Mapping
from dataclasses import dataclass
from typing import Optional

import torch

from . import _cabi, ops

_KEYS = ("loss", "confidence", "interaction", "weight")


@dataclass
class TrainConfig:
    """Same fields and defaults as the reference's ``TrainConfig`` (``dddm/training.py:16-29``)."""

    beta: float = 0.1
    lam: float = 1.0
    m: int = 8
    w_bias: float = 0.0
    lr: float = 2e-3
    epochs: int = 2000
    batch: int = 512
    device: str = "cpu"
    seed: int = 0
    use_wandb: bool = False
    wandb_project: str = "dddm"
    wandb_run_name: Optional[str] = None


class DeferredMetrics(Mapping):
    """The 4-key metrics mapping of the reference, backed by ONE device tensor.

    Values are read back (one packed 16-byte device->host copy, one synchronisation) the first
    time any key is accessed, instead of the reference's four blocking ``.cpu()`` reads per step
    (``dddm/training.py:87-92``).  ``.tensor`` exposes the device values for collective logging.
    """

    def __init__(self, packed: torch.Tensor):
        self.tensor = packed  # fp32 [4] on device: loss, conf, inter, weight
        self._host: dict | None = None

    def _materialise(self) -> dict:
        if self._host is None:
            self._host = dict(zip(_KEYS, (float(v) for v in self.tensor.tolist())))
        return self._host

    def __getitem__(self, key):
        return self._materialise()[key]

    def __iter__(self):
        return iter(_KEYS)

    def __len__(self):
        return len(_KEYS)


def _world(group) -> int:
    import torch.distributed as dist

    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def distributional_training_step(
    model: torch.nn.Module,
    x0: torch.Tensor,
    *,
    m: int,
    beta: float,
    lam: float,
    w_bias: float,
    t: Optional[torch.Tensor] = None,
    eps: Optional[torch.Tensor] = None,
    xi: Optional[torch.Tensor] = None,
    global_weight: Optional[bool] = None,
    group=None,
    sync_metrics: bool = True,
    fused_io: Optional[bool] = None,
    weight_sum: Optional[torch.Tensor] = None,
):
    """Generalized energy training loss (paper eqs. 12-14) — reference ``dddm/training.py:32-93``.

    Same positional/keyword structure, ``ValueError`` for ``m < 2``, RNG consumption order
    (``rand(B)`` unless ``t`` is given, ``randn_like(x0)``, ``randn(B, m, ...)``) and return value
    ``(loss, metrics)`` with ``metrics`` holding Python floats under the keys ``loss, confidence,
    interaction, weight``.  Keyword-only extensions (all optional):

    * ``eps``, ``xi``: pass the noise in instead of drawing it (parity tests);
    * ``global_weight`` / ``group``: under ``torch.distributed`` the logistic weight is the mean over
      the GLOBAL batch — one float all-reduce of sum_b w(t_b) issued before the backbone forward —
      so that a batch-sharded run equals the single-process global batch (the
      reference's loss is a product of two batch means).  Defaults to on when world_size > 1;
    * ``weight_sum``: sum_b w(t_b) over the GLOBAL batch, already all-reduced by the caller (the launcher does
      this outside its CUDA graphs); K4 and the all-reduce are then skipped.  Requires ``t``;
    * ``sync_metrics=False`` returns a :class:`DeferredMetrics` (no host synchronisation in the step);
    * ``fused_io`` (default: on when the model offers ``forward_cat``/``patch_size``, i.e. ``ddm_b200.backbones
      .DDDMDiT``, and ``x0`` is an image batch): K2c writes cat(x_t, xi) m-fold directly in the backbone's compute
      dtype and hands x0 over in patch-token order, and K1 reads the backbone's tokens without the unpatchify copy
      (SURVEY.md §8f-2).  Same loss (the score is invariant to a common permutation of D), same RNG order.

    Kernels: K4 (weight sum) -> K2 (marginal + m-fold expansion, written straight into the
    backbone's input) -> backbone (PyTorch) -> K1 (fused loss forward + backward).
    """
    if m < 2:
        raise ValueError("m must be >= 2 to form interaction pairs")
    if not x0.is_cuda:
        raise RuntimeError("ddm_b200.distributional_training_step runs on CUDA tensors only (no CPU fallback)")

    device, dtype, batch = x0.device, x0.dtype, x0.shape[0]
    if t is None:
        t = torch.rand(batch, device=device, dtype=dtype)
    if eps is None:
        eps = torch.randn_like(x0)
    if xi is None:
        xi = torch.randn((batch, m, *x0.shape[1:]), device=device, dtype=dtype)

    # logistic weight: per-rank sum now, (async) global sum while the backbone runs
    world = _world(group)
    use_global = (world > 1) if global_weight is None else (bool(global_weight) and world > 1)
    work = None
    if weight_sum is not None:
        w_sum = weight_sum.reshape(-1)[:1]
    else:
        _, w_sum = ops.sigmoid_weight_sum(t, float(w_bias))
        if use_global:
            import torch.distributed as dist

            work = dist.all_reduce(w_sum, op=dist.ReduceOp.SUM, group=group, async_op=True)
    weight_scale = 1.0 / (batch * (world if use_global else 1))

    t_rep = t.repeat_interleave(m)
    patch = int(getattr(model, "patch_size", 0) or 0)
    can_fuse = (hasattr(model, "forward_cat") and x0.dim() == 4 and patch >= 4 and patch % 4 == 0 and
                x0.shape[-1] % patch == 0 and x0.shape[-2] % patch == 0 and not x0.requires_grad)
    if fused_io and not can_fuse:
        raise ValueError("fused_io=True needs a backbone with forward_cat/patch_size and an image batch x0 [B,C,H,W]")
    x0_flat = None
    if can_fuse and fused_io is not False:
        wdtype = next(model.parameters()).dtype
        x6, x0_flat = ops.forward_marginal_concat(x0, t, eps, xi, wdtype == torch.bfloat16, patch)
        x0hat = model.forward_cat(x6, t_rep, tokens=True).reshape(batch, m, -1)  # patch-token order, no unpatchify
    else:
        _, xt_rep = ops.forward_marginal_expand(x0, t, eps, m, False)
        xi_flat = xi.reshape(batch * m, *x0.shape[1:])
        x0hat = model(xt_rep, t_rep, xi_flat)
        x0hat = x0hat.view(batch, m, *x0.shape[1:])

    if work is not None:
        work.wait()
    want_grad = torch.is_grad_enabled() and x0hat.requires_grad
    if x0.requires_grad and torch.is_grad_enabled():
        # gradient w.r.t. the data is only provided by the split kernels
        from .losses import generalized_energy_terms

        conf, inter = generalized_energy_terms(x0hat.reshape(batch, m, -1), x0.reshape(batch, -1), beta, lam)
        weight = (w_sum * weight_scale).reshape(()).to(dtype)
        loss = weight * (conf - (lam / (2.0 * (m - 1))) * inter)
        packed = torch.stack([loss.detach().float(), conf.detach().float(), inter.detach().float(), weight.float()])
    else:
        xh = x0hat.reshape(batch, m, -1)
        x0_2d = x0_flat if x0_flat is not None else x0.reshape(batch, -1)
        if xh.dtype != x0_2d.dtype:
            # a bf16 backbone under fp32 data: hand K1 the draws as they are (mixed entry point) instead of
            # up-converting them — one pass over xhat and one over the gradient less
            mixed = (xh.dtype == torch.bfloat16 and x0_2d.dtype == torch.float32 and
                     bool(_cabi.lib().dddm_energy_fused_bf16_x0f32_supported(m, xh.shape[2])))
            if not mixed:
                xh = xh.to(dtype)
        packed, _ = ops.energy_fused(xh, x0_2d, w_sum, weight_scale, float(beta), float(lam), want_grad)
        loss = packed[0].to(dtype)
        packed = packed.detach()

    metrics = DeferredMetrics(packed)
    if sync_metrics:
        metrics = dict(metrics._materialise())
    return loss, metrics
