"""PyTorch custom ops over the C ABI (``include/dddm_b200.h``).

torch is plumbing here: it owns the device memory and the stream; every op body is one call
into ``libdddm_b200.so`` with raw device pointers.  All ops are CUDA-only and raise if handed
CPU tensors — there is no eager fallback.  ``register_fake`` gives shape/dtype propagation so
that DDP / ``torch.compile`` tracing does not break, ``register_autograd`` wires the fused and
split energy-score gradients into autograd.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _cabi

_SUFFIX = {torch.float32: "f32", torch.bfloat16: "bf16"}


def _suffix(t: Tensor) -> str:
    try:
        return _SUFFIX[t.dtype]
    except KeyError:
        raise TypeError(f"ddm_b200 kernels support float32 and bfloat16, got {t.dtype}") from None


def _require_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ddm_b200 is CUDA-only (sm_100a kernels, no CPU fallback): got a tensor on "
                               f"{t.device}; move inputs to a CUDA device")


def _stream(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Tensor | None):
    return None if t is None else t.data_ptr()


_workspaces: dict = {}


def _workspace(ref: Tensor, B: int, m: int) -> Tensor:
    """Zero-initialised energy workspace, one per (device, stream): launches on one stream are ordered."""
    key = (ref.device.index, _stream(ref))
    need = _cabi.lib().dddm_energy_workspace_bytes(B, m)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 4096), dtype=torch.uint8, device=ref.device)
        _workspaces[key] = ws
    return ws


# --------------------------------------------------------------------------------------------
# K1: fused energy-score loss forward + backward  (dddm/training.py:77-85, dddm/losses.py:5-25)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("ddm_b200::energy_fused", mutates_args=())
def energy_fused(xhat: Tensor, x0: Tensor, weight: Tensor, weight_scale: float, beta: float, lam: float,
                 want_grad: bool) -> Tuple[Tensor, Tensor]:
    """Returns (out[4] = {loss, conf, inter, W} fp32, dloss/dxhat or an empty tensor)."""
    _require_cuda(xhat, x0, weight)
    if xhat.dim() != 3 or x0.dim() != 2 or x0.shape[0] != xhat.shape[0] or x0.shape[1] != xhat.shape[2]:
        raise ValueError(f"expected xhat [B,m,D] and x0 [B,D], got {tuple(xhat.shape)} and {tuple(x0.shape)}")
    sfx = _suffix(xhat)
    if x0.dtype != xhat.dtype:
        # mixed entry point: bf16 draws (as a bf16 backbone wrote them) against fp32 data, bf16 gradient out
        if not (xhat.dtype == torch.bfloat16 and x0.dtype == torch.float32):
            raise TypeError("xhat and x0 must have the same dtype (or bf16 xhat with fp32 x0)")
        if not _cabi.lib().dddm_energy_fused_bf16_x0f32_supported(xhat.shape[1], xhat.shape[2]):
            raise TypeError(f"bf16 xhat with fp32 x0 is not covered for m={xhat.shape[1]}, D={xhat.shape[2]}: "
                            "pass both in one dtype")
        sfx = "bf16_x0f32"
    xhat, x0 = xhat.contiguous(), x0.contiguous()
    weight = weight.reshape(-1)[:1].float().contiguous()
    B, m, D = xhat.shape
    out = torch.empty(4, dtype=torch.float32, device=xhat.device)
    grad = torch.empty_like(xhat) if want_grad else torch.empty(0, dtype=xhat.dtype, device=xhat.device)
    with torch.cuda.device(xhat.device):
        ws = _workspace(xhat, B, m)
        fn = getattr(_cabi.lib(), f"dddm_energy_fused_{sfx}")
        _cabi.check(fn(_ptr(xhat), _ptr(x0), _ptr(weight), float(weight_scale), _ptr(grad) if want_grad else None,
                       _ptr(out), _ptr(ws), B, m, D, float(beta), float(lam), _stream(xhat)))
    return out, grad


@energy_fused.register_fake
def _(xhat, x0, weight, weight_scale, beta, lam, want_grad):
    out = xhat.new_empty(4, dtype=torch.float32)
    return out, (torch.empty_like(xhat) if want_grad else xhat.new_empty(0))


@torch.library.custom_op("ddm_b200::scale_", mutates_args=("y",))
def scale_(y: Tensor, scale: Tensor) -> None:
    """y *= scale[0] unless scale[0] == 1 (tested on the device; no host synchronisation)."""
    _require_cuda(y, scale)
    sfx = _suffix(y)
    if not y.is_contiguous():
        raise ValueError("scale_ needs a contiguous tensor")
    scale = scale.reshape(-1)[:1].float().contiguous()
    with torch.cuda.device(y.device):
        fn = getattr(_cabi.lib(), f"dddm_scale_inplace_{sfx}")
        _cabi.check(fn(_ptr(y), _ptr(scale), y.numel(), _stream(y)))


def _energy_fused_setup(ctx, inputs, output):
    ctx.want_grad = inputs[6]
    ctx.save_for_backward(output[1])
    ctx.consumed = False


def _energy_fused_backward(ctx, g_out, g_grad):
    if not ctx.want_grad:
        raise RuntimeError("energy_fused was called with want_grad=False but its output is being differentiated")
    if ctx.consumed:
        raise RuntimeError("the fused energy-score gradient buffer was already consumed by a previous backward(); "
                           "use generalized_energy_terms (split kernels) when you need retain_graph=True")
    (grad,) = ctx.saved_tensors
    ctx.consumed = True
    # the kernel emitted dloss/dxhat assuming an upstream gradient of 1 for out[0] (the loss); only
    # out[0] is differentiable — conf/inter/W in out[1:] are logging values.
    scale_(grad, g_out[:1])
    return grad, None, None, None, None, None, None


energy_fused.register_autograd(_energy_fused_backward, setup_context=_energy_fused_setup)


# --------------------------------------------------------------------------------------------
# K1b: generalized_energy_terms forward / backward pair  (dddm/losses.py:5-25)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("ddm_b200::energy_terms_fwd", mutates_args=())
def energy_terms_fwd(xhat: Tensor, x0: Tensor, beta: float) -> Tuple[Tensor, Tensor]:
    """Returns (out[2] = {conf, inter} fp32, dist [B, m + m(m-1)/2] fp32 squared distances)."""
    _require_cuda(xhat, x0)
    if xhat.dim() != 3 or x0.dim() != 2 or x0.shape[0] != xhat.shape[0] or x0.shape[1] != xhat.shape[2]:
        raise ValueError(f"expected xhat [B,m,D] and x0 [B,D], got {tuple(xhat.shape)} and {tuple(x0.shape)}")
    sfx = _suffix(xhat)
    if x0.dtype != xhat.dtype:
        # mixed entry point: bf16 draws (as a bf16 backbone wrote them) against fp32 data, bf16 gradient out
        if not (xhat.dtype == torch.bfloat16 and x0.dtype == torch.float32):
            raise TypeError("xhat and x0 must have the same dtype (or bf16 xhat with fp32 x0)")
        if not _cabi.lib().dddm_energy_fused_bf16_x0f32_supported(xhat.shape[1], xhat.shape[2]):
            raise TypeError(f"bf16 xhat with fp32 x0 is not covered for m={xhat.shape[1]}, D={xhat.shape[2]}: "
                            "pass both in one dtype")
        sfx = "bf16_x0f32"
    xhat, x0 = xhat.contiguous(), x0.contiguous()
    B, m, D = xhat.shape
    out = torch.empty(2, dtype=torch.float32, device=xhat.device)
    dist = torch.empty((B, m + m * (m - 1) // 2), dtype=torch.float32, device=xhat.device)
    with torch.cuda.device(xhat.device):
        ws = _workspace(xhat, B, m)
        fn = getattr(_cabi.lib(), f"dddm_energy_terms_fwd_{sfx}")
        _cabi.check(fn(_ptr(xhat), _ptr(x0), _ptr(dist), _ptr(out), _ptr(ws), B, m, D, float(beta), _stream(xhat)))
    return out, dist


@energy_terms_fwd.register_fake
def _(xhat, x0, beta):
    B, m, _ = xhat.shape
    return xhat.new_empty(2, dtype=torch.float32), xhat.new_empty((B, m + m * (m - 1) // 2), dtype=torch.float32)


@torch.library.custom_op("ddm_b200::energy_terms_bwd", mutates_args=())
def energy_terms_bwd(xhat: Tensor, x0: Tensor, dist: Tensor, g_conf: Tensor, g_inter: Tensor, beta: float,
                     need_x0: bool) -> Tuple[Tensor, Tensor]:
    _require_cuda(xhat, x0, dist, g_conf, g_inter)
    sfx = _suffix(xhat)
    xhat, x0, dist = xhat.contiguous(), x0.contiguous(), dist.contiguous()
    g_conf = g_conf.reshape(-1)[:1].float().contiguous()
    g_inter = g_inter.reshape(-1)[:1].float().contiguous()
    B, m, D = xhat.shape
    gx = torch.empty_like(xhat)
    gx0 = torch.empty_like(x0) if need_x0 else torch.empty(0, dtype=x0.dtype, device=x0.device)
    with torch.cuda.device(xhat.device):
        fn = getattr(_cabi.lib(), f"dddm_energy_terms_bwd_{sfx}")
        _cabi.check(fn(_ptr(xhat), _ptr(x0), _ptr(dist), _ptr(g_conf), _ptr(g_inter), _ptr(gx),
                       _ptr(gx0) if need_x0 else None, B, m, D, float(beta), _stream(xhat)))
    return gx, gx0


@energy_terms_bwd.register_fake
def _(xhat, x0, dist, g_conf, g_inter, beta, need_x0):
    return torch.empty_like(xhat), (torch.empty_like(x0) if need_x0 else x0.new_empty(0))


def _energy_terms_setup(ctx, inputs, output):
    xhat, x0, beta = inputs
    ctx.beta = beta
    ctx.save_for_backward(xhat, x0, output[1])


def _energy_terms_backward(ctx, g_out, g_dist):
    xhat, x0, dist = ctx.saved_tensors
    need_x0 = ctx.needs_input_grad[1]
    gx, gx0 = energy_terms_bwd(xhat, x0, dist, g_out[0:1], g_out[1:2], ctx.beta, need_x0)
    return gx, (gx0 if need_x0 else None), None


energy_terms_fwd.register_autograd(_energy_terms_backward, setup_context=_energy_terms_setup)


# --------------------------------------------------------------------------------------------
# K2: forward marginal + m-fold expansion  (dddm/schedules.py:17-25, dddm/training.py:70)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("ddm_b200::forward_marginal_expand", mutates_args=())
def forward_marginal_expand(x0: Tensor, t: Tensor, eps: Tensor, m: int, want_xt: bool) -> Tuple[Tensor, Tensor]:
    """x0, eps [B, ...] same shape, t [B] fp32.  Returns (xt [B, ...] or empty, xt_rep [B*m, ...] or empty if m == 0)."""
    _require_cuda(x0, t, eps)
    sfx = _suffix(x0)
    if eps.shape != x0.shape or eps.dtype != x0.dtype:
        raise ValueError("eps must match x0 in shape and dtype")
    x0, eps = x0.contiguous(), eps.contiguous()
    t = t.reshape(-1).float().contiguous()
    B = x0.shape[0]
    if t.numel() != B:
        raise ValueError("t must have one entry per row of x0")
    D = x0.numel() // B if B else 0
    xt = torch.empty_like(x0) if want_xt else torch.empty(0, dtype=x0.dtype, device=x0.device)
    rep = (torch.empty((B * m, *x0.shape[1:]), dtype=x0.dtype, device=x0.device) if m > 0 else
           torch.empty(0, dtype=x0.dtype, device=x0.device))
    if not want_xt and m <= 0:
        raise ValueError("nothing to compute: want_xt is False and m == 0")
    with torch.cuda.device(x0.device):
        fn = getattr(_cabi.lib(), f"dddm_forward_marginal_expand_{sfx}")
        _cabi.check(fn(_ptr(x0), _ptr(t), _ptr(eps), _ptr(xt) if want_xt else None, _ptr(rep) if m > 0 else None, B,
                       max(m, 1), D, _stream(x0)))
    return xt, rep


@forward_marginal_expand.register_fake
def _(x0, t, eps, m, want_xt):
    B = x0.shape[0]
    xt = torch.empty_like(x0) if want_xt else x0.new_empty(0)
    rep = x0.new_empty((B * m, *x0.shape[1:])) if m > 0 else x0.new_empty(0)
    return xt, rep


def _fm_setup(ctx, inputs, output):
    x0, t, eps, m, want_xt = inputs
    ctx.m, ctx.want_xt = m, want_xt
    ctx.save_for_backward(x0, t, eps)


def _fm_backward(ctx, g_xt, g_rep):
    # rare path (the training step never differentiates through the marginal): plain tensor algebra
    x0, t, eps = ctx.saved_tensors
    B = x0.shape[0]
    g = torch.zeros_like(x0, dtype=torch.float32)
    if ctx.want_xt and g_xt is not None:
        g = g + g_xt.float()
    if ctx.m > 0 and g_rep is not None:
        g = g + g_rep.float().reshape(B, ctx.m, *x0.shape[1:]).sum(dim=1)
    tt = t.float().reshape(B, *([1] * (x0.dim() - 1)))
    g_x0 = ((1.0 - tt) * g).to(x0.dtype) if ctx.needs_input_grad[0] else None
    g_eps = (tt * g).to(eps.dtype) if ctx.needs_input_grad[2] else None
    g_t = ((eps.float() - x0.float()) * g).reshape(B, -1).sum(dim=1).to(t.dtype) if ctx.needs_input_grad[1] else None
    return g_x0, g_t, g_eps, None, None


forward_marginal_expand.register_autograd(_fm_backward, setup_context=_fm_setup)


@torch.library.custom_op("ddm_b200::forward_marginal_concat", mutates_args=())
def forward_marginal_concat(x0: Tensor, t: Tensor, eps: Tensor, xi: Tensor, out_bf16: bool,
                            patch: int) -> Tuple[Tensor, Tensor]:
    """K2c: x0, eps [B,C,H,W]; xi [B,m,C,H,W]; t [B].  Returns (x6 [B*m, 2C, H, W] = cat(x_t m-fold, xi) in fp32 or
    bf16, x0 in patch-token order [B, C*H*W] or an empty tensor when patch == 0).  Not differentiable (data and noise)."""
    _require_cuda(x0, t, eps, xi)
    sfx = _suffix(x0)
    if x0.dim() != 4 or eps.shape != x0.shape or xi.dim() != 5 or xi.shape[0] != x0.shape[0] or xi.shape[2:] != x0.shape[1:]:
        raise ValueError(f"expected x0/eps [B,C,H,W] and xi [B,m,C,H,W], got {tuple(x0.shape)}, {tuple(eps.shape)}, "
                         f"{tuple(xi.shape)}")
    if eps.dtype != x0.dtype or xi.dtype != x0.dtype:
        raise TypeError("x0, eps and xi must have the same dtype")
    x0, eps, xi = x0.contiguous(), eps.contiguous(), xi.contiguous()
    t = t.reshape(-1).float().contiguous()
    B, C, H, W = x0.shape
    m = xi.shape[1]
    if t.numel() != B:
        raise ValueError("t must have one entry per row of x0")
    out_dtype = torch.bfloat16 if (out_bf16 or x0.dtype == torch.bfloat16) else torch.float32
    x6 = torch.empty((B * m, 2 * C, H, W), dtype=out_dtype, device=x0.device)
    tok = torch.empty((B, C * H * W) if patch > 0 else (0,), dtype=x0.dtype, device=x0.device)
    with torch.cuda.device(x0.device):
        L = _cabi.lib()
        if sfx == "f32":
            _cabi.check(L.dddm_forward_marginal_concat_f32(_ptr(x0), _ptr(t), _ptr(eps), _ptr(xi), _ptr(x6),
                                                           1 if out_dtype == torch.bfloat16 else 0,
                                                           _ptr(tok) if patch > 0 else None, B, m, C, H, W, patch, _stream(x0)))
        else:
            _cabi.check(L.dddm_forward_marginal_concat_bf16(_ptr(x0), _ptr(t), _ptr(eps), _ptr(xi), _ptr(x6),
                                                            _ptr(tok) if patch > 0 else None, B, m, C, H, W, patch,
                                                            _stream(x0)))
    return x6, tok


@forward_marginal_concat.register_fake
def _(x0, t, eps, xi, out_bf16, patch):
    B, C, H, W = x0.shape
    dt = torch.bfloat16 if (out_bf16 or x0.dtype == torch.bfloat16) else torch.float32
    return (x0.new_empty((B * xi.shape[1], 2 * C, H, W), dtype=dt),
            x0.new_empty((B, C * H * W) if patch > 0 else (0,)))


# --------------------------------------------------------------------------------------------
# K4: logistic weight and its batch sum  (dddm/losses.py:28-35, dddm/training.py:84)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("ddm_b200::sigmoid_weight_sum", mutates_args=())
def sigmoid_weight_sum(t: Tensor, bias: float) -> Tuple[Tensor, Tensor]:
    """t [B] -> (w [B] fp32, sum_b w [1] fp32)."""
    _require_cuda(t)
    t = t.reshape(-1).float().contiguous()
    B = t.numel()
    w = torch.empty(B, dtype=torch.float32, device=t.device)
    w_sum = torch.empty(1, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _cabi.check(_cabi.lib().dddm_sigmoid_weight_sum_f32(_ptr(t), float(bias), _ptr(w), _ptr(w_sum), B, _stream(t)))
    return w, w_sum


@sigmoid_weight_sum.register_fake
def _(t, bias):
    return t.new_empty(t.numel(), dtype=torch.float32), t.new_empty(1, dtype=torch.float32)


# --------------------------------------------------------------------------------------------
# K3: Gaussian bridge + Algorithm-2 update  (dddm/schedules.py:28-78, dddm/sampling.py:29-31)
# --------------------------------------------------------------------------------------------
def _bridge_call(x_out, x, xhat0, z, s, t, eps_churn, mu_out, std_out):
    sfx = _suffix(x)
    N = x.shape[0]
    D = x.numel() // N if N else 0
    vec = int(s.numel() > 1)
    with torch.cuda.device(x.device):
        fn = getattr(_cabi.lib(), f"dddm_bridge_step_{sfx}")
        _cabi.check(fn(_ptr(x_out), _ptr(x), _ptr(xhat0), _ptr(z), _ptr(s), _ptr(t), vec, float(eps_churn),
                       _ptr(mu_out), _ptr(std_out), N, D, _stream(x)))


def _prep_st(s: Tensor, t: Tensor, N: int, device) -> Tuple[Tensor, Tensor]:
    s = s.reshape(-1).float().contiguous()
    t = t.reshape(-1).float().contiguous()
    if s.numel() != t.numel():
        n = max(s.numel(), t.numel())
        s, t = s.expand(n).contiguous(), t.expand(n).contiguous()
    if s.numel() not in (1, N):
        raise ValueError(f"s and t must be scalars or have one entry per sample ({N}), got {s.numel()}")
    return s, t


@torch.library.custom_op("ddm_b200::bridge_step", mutates_args=())
def bridge_step(x: Tensor, xhat0: Tensor, z: Tensor, s: Tensor, t: Tensor, eps_churn: float) -> Tensor:
    """x_next = mu(s, t, xhat0, x) + std(s, t) * z   (one Algorithm-2 update)."""
    _require_cuda(x, xhat0, z, s, t)
    if xhat0.shape != x.shape or z.shape != x.shape or xhat0.dtype != x.dtype or z.dtype != x.dtype:
        raise ValueError("x, xhat0 and z must match in shape and dtype")
    x, xhat0, z = x.contiguous(), xhat0.contiguous(), z.contiguous()
    s, t = _prep_st(s, t, x.shape[0], x.device)
    out = torch.empty_like(x)
    _bridge_call(out, x, xhat0, z, s, t, eps_churn, None, None)
    return out


@bridge_step.register_fake
def _(x, xhat0, z, s, t, eps_churn):
    return torch.empty_like(x)


def philox_increment(numel: int, device) -> int:
    """How far one ``torch.randn_like`` over ``numel`` elements advances the CUDA generator's Philox offset on
    ``device`` (ATen's ``calc_execution_policy``)."""
    with torch.cuda.device(device):
        return int(_cabi.lib().dddm_philox_increment(int(numel)))


def bridge_step_philox_(x: Tensor, xhat0: Tensor, xi_next: Optional[Tensor], s: Tensor, t: Tensor, eps_churn: float, *,
                        philox: Optional[Tensor] = None, seed: int = 0, offset_z: int = 0, offset_xi: int = 0) -> Tensor:
    """In-place Algorithm-2 update with the step's Gaussian draws fused in (``dddm/sampling.py:27-31``):
    ``x <- mu(s, t, xhat0, x) + std(s, t) * z`` with ``z = randn_like(x)`` generated in registers at Philox offset
    ``offset_z``, and — when ``xi_next`` is given — the next step's ``xi = randn_like(x)`` written at ``offset_xi``.
    Bit-identical to drawing both with ``torch.randn_like`` at those generator offsets.  ``philox``: optional device
    int64[3] ``{seed, offset_z, offset_xi}`` read by the kernel instead of the keyword values (CUDA-graph replay)."""
    _require_cuda(x, xhat0, s, t)
    if xhat0.shape != x.shape or xhat0.dtype != x.dtype:
        raise ValueError("x and xhat0 must match in shape and dtype")
    if not x.is_contiguous():
        raise ValueError("x must be contiguous (it is updated in place)")
    if xi_next is not None and (xi_next.shape != x.shape or xi_next.dtype != x.dtype or not xi_next.is_contiguous()):
        raise ValueError("xi_next must be a contiguous tensor like x")
    if philox is not None and (philox.dtype != torch.int64 or philox.numel() < 3 or philox.device != x.device):
        raise ValueError("philox must be a device int64 tensor {seed, offset_z, offset_xi}")
    xhat0 = xhat0.contiguous()
    s, t = _prep_st(s, t, x.shape[0], x.device)
    if s.numel() != 1:
        raise ValueError("the fused-noise update takes scalar times (one Algorithm-2 step)")
    N = x.shape[0]
    D = x.numel() // N if N else 0
    mask = (1 << 64) - 1
    with torch.cuda.device(x.device):
        fn = getattr(_cabi.lib(), f"dddm_bridge_step_philox_{_suffix(x)}")
        _cabi.check(fn(_ptr(x), _ptr(x), _ptr(xhat0), _ptr(xi_next), _ptr(s), _ptr(t), float(eps_churn), _ptr(philox),
                       int(seed) & mask, int(offset_z) & mask, int(offset_xi) & mask, N, D, _stream(x)))
    return x


@torch.library.custom_op("ddm_b200::bridge_mu_sigma", mutates_args=())
def bridge_mu_sigma(s: Tensor, t: Tensor, x0: Tensor, xt: Tensor, eps_churn: float) -> Tuple[Tensor, Tensor]:
    """(mu [N, ...], std [1 or N] fp32) of gaussian_bridge_mu_sigma."""
    _require_cuda(s, t, x0, xt)
    if x0.shape != xt.shape or x0.dtype != xt.dtype:
        raise ValueError("x0 and xt must match in shape and dtype")
    x0, xt = x0.contiguous(), xt.contiguous()
    s, t = _prep_st(s, t, xt.shape[0], xt.device)
    mu = torch.empty_like(xt)
    std = torch.empty(s.numel(), dtype=torch.float32, device=xt.device)
    _bridge_call(None, xt, x0, None, s, t, eps_churn, mu, std)
    return mu, std


@bridge_mu_sigma.register_fake
def _(s, t, x0, xt, eps_churn):
    return torch.empty_like(xt), xt.new_empty(max(s.numel(), t.numel()), dtype=torch.float32)


# --------------------------------------------------------------------------------------------
# Backbone helpers: LayerNorm forward/backward and column sums for the DiT step
# --------------------------------------------------------------------------------------------
def _scratch(ref: Tensor, C: int) -> Tensor:
    return torch.empty(_cabi.lib().dddm_backbone_scratch_bytes(C), dtype=torch.uint8, device=ref.device)


def layer_norm_supported(x: Tensor, weight: Tensor, bias: Tensor) -> bool:
    C = x.shape[-1]
    return (x.is_cuda and x.dtype in _SUFFIX and weight is not None and bias is not None and weight.dtype == x.dtype
            and bias.dtype == x.dtype and C % 4 == 0 and 4 <= C <= 1024 and (C * x.element_size()) % 16 == 0)


@torch.library.custom_op("ddm_b200::layer_norm", mutates_args=())
def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """LayerNorm over the last dimension: returns (y like x, mean [N] fp32, rstd [N] fp32)."""
    _require_cuda(x, weight, bias)
    sfx = _suffix(x)
    x = x.contiguous()
    C = x.shape[-1]
    N = x.numel() // C
    y = torch.empty_like(x)
    mean = torch.empty(N, dtype=torch.float32, device=x.device)
    rstd = torch.empty(N, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        fn = getattr(_cabi.lib(), f"dddm_layer_norm_fwd_{sfx}")
        _cabi.check(fn(_ptr(x), _ptr(weight.contiguous()), _ptr(bias.contiguous()), _ptr(y), _ptr(mean), _ptr(rstd), N, C,
                       float(eps), _stream(x)))
    return y, mean, rstd


@layer_norm.register_fake
def _(x, weight, bias, eps):
    N = x.numel() // x.shape[-1]
    return torch.empty_like(x), x.new_empty(N, dtype=torch.float32), x.new_empty(N, dtype=torch.float32)


@torch.library.custom_op("ddm_b200::layer_norm_bwd", mutates_args=())
def layer_norm_bwd(dy: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, weight: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    _require_cuda(dy, x, mean, rstd, weight)
    sfx = _suffix(x)
    dy, x = dy.contiguous(), x.contiguous()
    C = x.shape[-1]
    N = x.numel() // C
    dx = torch.empty_like(x)
    dw, db = torch.empty_like(weight), torch.empty_like(weight)
    with torch.cuda.device(x.device):
        scratch = _scratch(x, C)
        fn = getattr(_cabi.lib(), f"dddm_layer_norm_bwd_{sfx}")
        _cabi.check(fn(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(weight.contiguous()), _ptr(dx), _ptr(dw), _ptr(db),
                       _ptr(scratch), scratch.numel(), N, C, _stream(x)))
    return dx, dw, db


@layer_norm_bwd.register_fake
def _(dy, x, mean, rstd, weight):
    return torch.empty_like(x), torch.empty_like(weight), torch.empty_like(weight)


def _ln_setup(ctx, inputs, output):
    x, weight, bias, eps = inputs
    ctx.save_for_backward(x, output[1], output[2], weight)


def _ln_backward(ctx, gy, gmean, grstd):
    x, mean, rstd, weight = ctx.saved_tensors
    dx, dw, db = layer_norm_bwd(gy, x, mean, rstd, weight)
    return dx, dw, db, None


layer_norm.register_autograd(_ln_backward, setup_context=_ln_setup)


def colsum_supported(a: Tensor) -> bool:
    C = a.shape[-1]
    return a.is_cuda and a.dtype in _SUFFIX and a.dim() == 2 and C % 4 == 0 and (C * a.element_size()) % 16 == 0 and a.shape[0] > 0


@torch.library.custom_op("ddm_b200::colsum", mutates_args=())
def colsum(a: Tensor) -> Tensor:
    """out[c] = sum_n a[n, c] (fp32 accumulation, fixed summation order), same dtype as ``a``."""
    _require_cuda(a)
    sfx = _suffix(a)
    a = a.contiguous()
    N, C = a.shape
    out = torch.empty(C, dtype=a.dtype, device=a.device)
    with torch.cuda.device(a.device):
        scratch = _scratch(a, C)
        fn = getattr(_cabi.lib(), f"dddm_colsum_{sfx}")
        _cabi.check(fn(_ptr(a), _ptr(out), _ptr(scratch), scratch.numel(), N, C, _stream(a)))
    return out


@colsum.register_fake
def _(a):
    return a.new_empty(a.shape[1])


# --------------------------------------------------------------------------------------------
# Evaluation: pairwise RBF kernel sums for rbf_mmd2  (dddm/metrics.py:140-163)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("ddm_b200::row_sqnorm", mutates_args=())
def row_sqnorm(x: Tensor) -> Tensor:
    _require_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2:
        raise TypeError("row_sqnorm expects a 2-D float32 tensor")
    x = x.contiguous()
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().dddm_row_sqnorm_f32(_ptr(x), _ptr(out), x.shape[0], x.shape[1], _stream(x)))
    return out


@row_sqnorm.register_fake
def _(x):
    return x.new_empty(x.shape[0])


@torch.library.custom_op("ddm_b200::rbf_kernel_sum", mutates_args=())
def rbf_kernel_sum(gram: Tensor, a2: Tensor, b2: Tensor, gamma: float, diag_shift: int, skip_diag: bool) -> Tensor:
    """sum over the tile of exp(-gamma (a2_r + b2_c - 2 G_rc)), skipping r + diag_shift == c when skip_diag; fp64 [1]."""
    _require_cuda(gram, a2, b2)
    if gram.dtype != torch.float32 or gram.dim() != 2 or gram.stride(1) != 1:
        raise TypeError("rbf_kernel_sum expects a row-major float32 Gram tile")
    rows, cols = gram.shape
    a2, b2 = a2.float().contiguous(), b2.float().contiguous()
    out = torch.empty(1, dtype=torch.float64, device=gram.device)
    with torch.cuda.device(gram.device):
        L = _cabi.lib()
        nbytes = L.dddm_rbf_scratch_bytes(rows, cols)
        scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=gram.device)
        _cabi.check(L.dddm_rbf_kernel_sum_f32(_ptr(gram), gram.stride(0), _ptr(a2), _ptr(b2), rows, cols, float(gamma),
                                              int(diag_shift), int(bool(skip_diag)), _ptr(scratch), nbytes, _ptr(out),
                                              _stream(gram)))
    return out


@rbf_kernel_sum.register_fake
def _(gram, a2, b2, gamma, diag_shift, skip_diag):
    return gram.new_empty(1, dtype=torch.float64)


@torch.library.custom_op("ddm_b200::rbf_split_bf16", mutates_args=())
def rbf_split_bf16(x: Tensor) -> Tuple[Tensor, Tensor]:
    """x fp32 [n, D] -> (hi, lo) bf16 [n, Dp] with x ~= hi + lo (Dp = D rounded up to 64, zero padded)."""
    _require_cuda(x)
    if x.dtype != torch.float32 or x.dim() != 2:
        raise TypeError("rbf_split_bf16 expects a 2-D float32 tensor")
    x = x.contiguous()
    n, D = x.shape
    L = _cabi.lib()
    Dp = int(L.dddm_rbf_tc_padded_cols(D))
    hi = torch.empty(n, Dp, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(n, Dp, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _cabi.check(L.dddm_rbf_split_bf16(_ptr(x), _ptr(hi), _ptr(lo), n, D, _stream(x)))
    return hi, lo


@rbf_split_bf16.register_fake
def _(x):
    Dp = (x.shape[1] + 63) // 64 * 64
    return x.new_empty(x.shape[0], Dp, dtype=torch.bfloat16), x.new_empty(x.shape[0], Dp, dtype=torch.bfloat16)


@torch.library.custom_op("ddm_b200::rbf_kernel_sum_tc", mutates_args=())
def rbf_kernel_sum_tc(a_hi: Tensor, a_lo: Tensor, b_hi: Tensor, b_lo: Tensor, a2: Tensor, b2: Tensor, D: int, gamma: float,
                      symmetric: bool) -> Tensor:
    """sum_{i,j} w_ij exp(-gamma (a2_i + b2_j - 2 a_i.b_j)) from the bf16 hi/lo splits, fused on the tensor cores; fp64 [1]."""
    _require_cuda(a_hi, a_lo, b_hi, b_lo, a2, b2)
    for t in (a_hi, a_lo, b_hi, b_lo):
        if t.dtype != torch.bfloat16 or t.dim() != 2 or not t.is_contiguous():
            raise TypeError("rbf_kernel_sum_tc expects contiguous 2-D bfloat16 splits (ops.rbf_split_bf16)")
    a2, b2 = a2.float().contiguous(), b2.float().contiguous()
    out = torch.empty(1, dtype=torch.float64, device=a_hi.device)
    with torch.cuda.device(a_hi.device):
        L = _cabi.lib()
        nbytes = L.dddm_rbf_tc_scratch_bytes()
        scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=a_hi.device)
        _cabi.check(L.dddm_rbf_kernel_sum_tc(_ptr(a_hi), _ptr(a_lo), _ptr(b_hi), _ptr(b_lo), _ptr(a2), _ptr(b2), a_hi.shape[0],
                                             b_hi.shape[0], int(D), float(gamma), int(bool(symmetric)), _ptr(scratch), nbytes,
                                             _ptr(out), _stream(a_hi)))
    return out


@rbf_kernel_sum_tc.register_fake
def _(a_hi, a_lo, b_hi, b_lo, a2, b2, D, gamma, symmetric):
    return a_hi.new_empty(1, dtype=torch.float64)
