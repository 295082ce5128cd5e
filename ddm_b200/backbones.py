"""PyTorch denoiser backbones x0hat = f(x_t, t, xi) used by the benchmarks and the launcher.

The backbones are OUT of the hot-path scope (BASELINE.json: "the MLP/DiT backbones stay in
PyTorch"); they exist here only because ``/root/reference`` is not present on the GPU box and the
DiT / sampler throughput configs need a model of the reference's architecture with random
weights.  Same architecture, parameter names and shapes as the reference's ``dddm/model.py``
(``DDDMMLP`` :41-67, ``DDDMDiT`` :183-244 — default DiT = 14,523,312 parameters), so a reference
checkpoint's ``state_dict`` loads.  Same math, arranged for throughput (measured with
``tools/profile_dit_step.py``, 65 536 tokens per step at BASELINE config 4):

* attention is ``scaled_dot_product_attention`` instead of the reference's explicit softmax(qk^T)v;
  q, k, v come from three GEMMs over row-slices of the single ``qkv`` weight, so SDPA consumes
  ``[B, n, h, d]`` views and its backward needs no unbind/stack/contiguous copies (7 ms / step);
* LayerNorm applies its affine outside the normalisation kernel: ATen's gamma/beta backward kernel
  took 0.56 ms per LayerNorm at this shape (9.6 ms / step, the largest single item); two plain
  reductions replace it;
* the time embedding is always evaluated in fp32 (a bf16 ``t`` has 8 bits of resolution);
* on CUDA, LayerNorm forward/backward and the bias-gradient column sums of every Linear run on this repo's
  kernels (``csrc/backbone_ops.cu``; ``USE_CUDA_KERNELS = False`` switches back to the plain ATen composition,
  which is also what runs on CPU).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


USE_CUDA_KERNELS = True  # LayerNorm / bias-gradient kernels of csrc/backbone_ops.cu on CUDA tensors (training only)


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b with the bias gradient reduced by ``ops.colsum`` (fixed-order column sums)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        from . import ops

        x, weight = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        gx = (g2 @ weight).view(x.shape) if ctx.needs_input_grad[0] else None
        gw = g2.t() @ x.reshape(-1, x.shape[-1]) if ctx.needs_input_grad[1] else None
        gb = None
        if ctx.needs_input_grad[2]:
            g2c = g2.contiguous()
            gb = ops.colsum(g2c) if ops.colsum_supported(g2c) else g2c.sum(dim=0)
        return gx, gw, gb


def _linear(x, weight, bias):
    if USE_CUDA_KERNELS and x.is_cuda and bias is not None and torch.is_grad_enabled() and x.dtype == weight.dtype:
        return _LinearFn.apply(x, weight, bias)
    return F.linear(x, weight, bias)


class _QKVFn(torch.autograd.Function):
    """q, k, v = x W_q^T + b_q, ... from row-slices of ONE packed weight (state-dict layout of the reference), as three
    GEMMs so that each of q/k/v is a dense [B, n, C] tensor SDPA can view as [B, n, h, d] without copies.  Backward
    accumulates dX over the three branches inside the GEMM epilogue (addmm, beta = 1) instead of two extra passes, and
    writes dW / db straight into the three row-slices of the packed gradients."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        c = x.shape[-1]
        ctx.save_for_backward(x, weight)
        return tuple(F.linear(x, weight[i * c:(i + 1) * c], bias[i * c:(i + 1) * c]) for i in range(3))

    @staticmethod
    def backward(ctx, gq, gk, gv):
        from . import ops

        x, weight = ctx.saved_tensors
        c = x.shape[-1]
        x2 = x.reshape(-1, c)
        gs = [g.reshape(-1, c).contiguous() for g in (gq, gk, gv)]
        gx = gs[0] @ weight[:c]
        gx.addmm_(gs[1], weight[c:2 * c])
        gx.addmm_(gs[2], weight[2 * c:])
        gw = torch.empty_like(weight)
        gb = torch.empty(3 * c, dtype=weight.dtype, device=weight.device)
        for i, g in enumerate(gs):
            torch.mm(g.t(), x2, out=gw[i * c:(i + 1) * c])
            gb[i * c:(i + 1) * c] = ops.colsum(g) if ops.colsum_supported(g) else g.sum(dim=0)
        return gx.view(x.shape), gw, gb


class _Linear(nn.Linear):
    def forward(self, x):
        return _linear(x, self.weight, self.bias)


class _FourierTime(nn.Module):
    """sin/cos of 2*pi*k*t for k = 1..n (reference ``TimeFeat``); ``freq`` is a frozen parameter."""

    def __init__(self, n: int):
        super().__init__()
        self.freq = nn.Parameter(torch.arange(1, n + 1, dtype=torch.float32), requires_grad=False)

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        ang = (2.0 * math.pi) * t[:, None] * self.freq[None, :]
        return torch.cat((ang.sin(), ang.cos()), dim=-1)


class DDDMMLP(nn.Module):
    """Toy 2-D denoiser: [x_t (2), xi (2), time features] -> x0hat (2); 4 hidden SiLU layers."""

    def __init__(self, time_dim: int = 32, hidden: int = 128):
        super().__init__()
        self.tfeat = _FourierTime(time_dim // 2)
        widths = [4 + time_dim] + [hidden] * 4
        layers: list[nn.Module] = []
        for a, b in zip(widths[:-1], widths[1:]):
            layers += [nn.Linear(a, b), nn.SiLU()]
        layers.append(nn.Linear(hidden, 2))
        self.net = nn.Sequential(*layers)

    def forward(self, xt, t, xi):
        return self.net(torch.cat((xt, xi, self.tfeat(t)), dim=-1))


def _sinusoidal(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    k = torch.arange(half, device=t.device, dtype=t.dtype)
    freqs = torch.exp(k * (-math.log(max_period) / max(half - 1, 1)))
    ang = t.reshape(-1, 1) * freqs
    emb = torch.cat((ang.sin(), ang.cos()), dim=-1)
    return F.pad(emb, (0, 1)) if dim % 2 else emb


class _Proj(nn.Module):
    """Holder so parameter names read ``<name>.proj.*`` like the reference's PatchEmbed / PatchUnembed."""

    def __init__(self, proj: nn.Module):
        super().__init__()
        self.proj = proj


class _LayerNorm(nn.LayerNorm):
    """LayerNorm with the affine applied as a separate fused multiply-add (same parameters, same math)."""

    def forward(self, x):
        if not torch.is_grad_enabled():  # inference (sampler): the single fused ATen kernel is the fastest form
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        if USE_CUDA_KERNELS and x.is_cuda:
            from . import ops

            if ops.layer_norm_supported(x, self.weight, self.bias):
                return ops.layer_norm(x, self.weight, self.bias, self.eps)[0]
        xn = F.layer_norm(x, self.normalized_shape, None, None, self.eps)
        return torch.addcmul(self.bias, xn, self.weight)


def _sdpa(q, k, v):
    """Short sequences (64 tokens here): the memory-efficient backend is ~1.7x faster than cuDNN's flash kernels on
    B200 for fwd+bwd at [1024, 6, 64, 64] and returns dq/dk/dv in the layout of its inputs (no re-layout copies)."""
    if q.is_cuda and q.shape[-2] <= 256 and q.dtype in (torch.bfloat16, torch.float16):
        from torch.nn.attention import SDPBackend, sdpa_kernel

        with sdpa_kernel([SDPBackend.EFFICIENT_ATTENTION, SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION,
                          SDPBackend.MATH], set_priority=True):
            return F.scaled_dot_product_attention(q, k, v)
    return F.scaled_dot_product_attention(q, k, v)


class _Attention(nn.Module):
    def __init__(self, dim: int, heads: int):
        super().__init__()
        if dim % heads:
            raise ValueError("dim must be divisible by num_heads")
        self.heads = heads
        self.qkv = _Linear(dim, 3 * dim)
        self.proj = _Linear(dim, dim)

    def forward(self, x):
        b, n, c = x.shape
        if not torch.is_grad_enabled():  # inference: one GEMM, q/k/v as strided views
            q, k, v = self.qkv(x).view(b, n, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
            return self.proj(_sdpa(q, k, v).transpose(1, 2).reshape(b, n, c))
        w, bias = self.qkv.weight, self.qkv.bias
        if USE_CUDA_KERNELS and x.is_cuda and x.dtype == w.dtype:
            qkv = _QKVFn.apply(x, w, bias)
        else:
            qkv = tuple(F.linear(x, w[i * c:(i + 1) * c], bias[i * c:(i + 1) * c]) for i in range(3))
        q, k, v = (t.view(b, n, self.heads, c // self.heads).transpose(1, 2) for t in qkv)
        return self.proj(_sdpa(q, k, v).transpose(1, 2).reshape(b, n, c))


class _MLP(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.net = nn.Sequential(_Linear(dim, hidden), nn.GELU(), _Linear(hidden, dim))

    def forward(self, x):
        return self.net(x)


class _Block(nn.Module):
    def __init__(self, dim: int, heads: int, mlp_ratio: float):
        super().__init__()
        self.norm1 = _LayerNorm(dim)
        self.attn = _Attention(dim, heads)
        self.norm2 = _LayerNorm(dim)
        self.ff = _MLP(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.ff(self.norm2(x))


class DDDMDiT(nn.Module):
    """DiT-S/4-style image denoiser: cat(x_t, xi) -> 4x4 patches -> ``depth`` pre-norm blocks -> unpatchify."""

    def __init__(self, img_size: int = 32, patch_size: int = 4, in_channels: int = 6, out_channels: int = 3,
                 embed_dim: int = 384, depth: int = 8, num_heads: int = 6, time_embed_dim: int = 256,
                 mlp_ratio: float = 4.0):
        super().__init__()
        if img_size % patch_size:
            raise ValueError("Image size must be divisible by patch size")
        self.img_size, self.patch_size, self.out_channels = img_size, patch_size, out_channels
        self.embed_dim, self.time_embed_dim = embed_dim, time_embed_dim
        self.grid = img_size // patch_size
        self.patch_embed = _Proj(nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size, stride=patch_size))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.grid * self.grid, embed_dim))
        self.time_mlp = nn.Sequential(_Linear(time_embed_dim, embed_dim), nn.SiLU(), _Linear(embed_dim, embed_dim))
        self.blocks = nn.ModuleList(_Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth))
        self.norm = _LayerNorm(embed_dim)
        self.unembed = _Proj(_Linear(embed_dim, out_channels * patch_size * patch_size))
        nn.init.trunc_normal_(self.pos_embed, std=0.02)

    def forward(self, xt, t, xi):
        if xt.shape != xi.shape:
            raise ValueError("xt and xi must have the same shape")
        if xt.dim() != 4:
            raise ValueError("Expecting image tensors with shape [B, C, H, W]")
        return self.forward_cat(torch.cat((xt, xi), dim=1), t)

    def forward_cat(self, x6, t, tokens: bool = False):
        """Same as ``forward`` with the channel concat cat(x_t, xi) already formed ([B, 6, H, W]); the
        training step's K2c kernel writes that tensor directly.  ``tokens=True`` returns
        PatchUnembed's projection ``[B, (H/p)(W/p), C*p*p]`` without the unpatchify permute/copy: the energy
        score only needs x0 in the same order (K2c provides it)."""
        wdtype = self.patch_embed.proj.weight.dtype
        # .contiguous(): the conv output viewed as tokens is a permuted-stride tensor; left as is, the whole residual
        # stream inherits those strides and every LayerNorm / Linear pays a 50 MB re-layout copy (34 per step)
        emb = self.patch_embed.proj(x6.to(wdtype)).flatten(2).transpose(1, 2).contiguous()
        temb = self.time_mlp(_sinusoidal(t.reshape(-1).float(), self.time_embed_dim).to(wdtype))
        h = emb + temb[:, None, :] + self.pos_embed
        for blk in self.blocks:
            h = blk(h)
        y = self.unembed.proj(self.norm(h))
        if tokens:
            return y
        g, p, c = self.grid, self.patch_size, self.out_channels
        y = y.view(-1, g, g, c, p, p).permute(0, 3, 1, 4, 2, 5)
        return y.reshape(-1, c, self.img_size, self.img_size)
