"""Drop-in for ``dddm/metrics.py::rbf_mmd2`` (evaluation side of the path)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

_ROW_CHUNK = 8192   # rows of a rectangular Gram tile: 8192 x 10 000 fp32 = 328 MB
_SYM_CHUNK = 2048   # rows of a trapezoidal tile of a symmetric term


def _kernel_sum(a: torch.Tensor, b: torch.Tensor, a2: torch.Tensor, b2: torch.Tensor, gamma: float) -> torch.Tensor:
    """sum_{i,j} k(a_i, b_j) over the full rectangle (the xy term, metrics.py:161)."""
    total = torch.zeros(1, dtype=torch.float64, device=a.device)
    for r0 in range(0, a.shape[0], _ROW_CHUNK):
        gram = a[r0:r0 + _ROW_CHUNK] @ b.t()  # library GEMM (cuBLAS), fp32 like the reference's `a @ b.T` (metrics.py:146)
        total += ops.rbf_kernel_sum(gram, a2[r0:r0 + _ROW_CHUNK], b2, gamma, 0, False)
    return total


def _kernel_sum_offdiag(a: torch.Tensor, a2: torch.Tensor, gamma: float) -> torch.Tensor:
    """sum_{i != j} k(a_i, a_j) (the xx / yy terms, metrics.py:157-160).  The matrix is symmetric: only the
    trapezoid right of each diagonal block is formed (0.6x the GEMM work at n = 10 000), counted twice."""
    n = a.shape[0]
    total = torch.zeros(1, dtype=torch.float64, device=a.device)
    for r0 in range(0, n, _SYM_CHUNK):
        r1 = min(r0 + _SYM_CHUNK, n)
        gram = a[r0:r1] @ a[r0:].t()  # [r1 - r0, n - r0]
        total += ops.rbf_kernel_sum(gram[:, :r1 - r0], a2[r0:r1], a2[r0:r1], gamma, 0, True)
        if r1 < n:
            total += 2.0 * ops.rbf_kernel_sum(gram[:, r1 - r0:], a2[r0:r1], a2[r1:], gamma, 0, False)
    return total


# Below this many multiply-adds per term the library-GEMM tile path is used: the fused tensor-core kernel splits the
# inputs into two bf16 terms (2^-17 relative per product, random sign) — invisible in a mean over millions of pairs,
# but a handful of pairs would see it at the 1e-4 level for sigma ~ 1.
_TC_MIN_WORK = 1 << 27


def _rbf_mmd2_tc(xf: torch.Tensor, yf: torch.Tensor, gamma: float) -> torch.Tensor:
    """The three terms of metrics.py:157-162 through the fused tcgen05 kernel (no Gram matrix in memory)."""
    n, m, D = xf.shape[0], yf.shape[0], xf.shape[1]
    x2, y2 = ops.row_sqnorm(xf), ops.row_sqnorm(yf)
    xh, xl = ops.rbf_split_bf16(xf)
    yh, yl = ops.rbf_split_bf16(yf)
    kxx = ops.rbf_kernel_sum_tc(xh, xl, xh, xl, x2, x2, D, gamma, True) / (n * (n - 1))
    kyy = ops.rbf_kernel_sum_tc(yh, yl, yh, yl, y2, y2, D, gamma, True) / (m * (m - 1))
    kxy = ops.rbf_kernel_sum_tc(xh, xl, yh, yl, x2, y2, D, gamma, False) / (n * m)
    return kxx + kyy - 2.0 * kxy


@torch.no_grad()
def rbf_mmd2(x: torch.Tensor, y: torch.Tensor, sigma: float = 1.0, *, allow_tf32: Optional[bool] = None,
             method: Optional[str] = None) -> torch.Tensor:
    """Unbiased MMD^2 with an RBF kernel, sigma fixed — reference ``dddm/metrics.py:140-163``.

    Same signature, ``ValueError`` for fewer than two samples per set, 0-dim result in the input dtype.  The three
    n x n terms are never materialised beyond one Gram tile each: GEMM tile -> one fused pass (distance, exp,
    diagonal mask, sum).  CUDA-only; not differentiable (the reference only calls it on detached samples,
    ``run_example.py:101``).

    Evaluation-sized sets (n * m * D >= 2^27, e.g. the 10 000 x 3072 of configs/cifar10_dit.yaml) take the FUSED path:
    one persistent tcgen05 kernel per term accumulates the Gram tile in tensor memory (bf16 hi/lo split of the fp32
    inputs, fp32 accumulation) and its epilogue reads it from there — no n x n matrix ever reaches HBM
    (``csrc/metrics_tc.cu``).  ``method='tc'`` / ``'gemm'`` force either path.

    ``allow_tf32`` (keyword-only extension, library-GEMM path): ``None`` keeps the process-wide ``torch.backends.cuda.matmul.allow_tf32``
    (PyTorch's default False = the reference's fp32 Gram); ``True`` runs the Gram tiles on the TF32 tensor cores —
    10x faster at n = 10 000, D = 3072 with the result unchanged to 6 digits there, but the Gram form's cancellation
    then carries 2^-11 instead of 2^-24 of ||x||^2, so it is opt-in.
    """
    if not (x.is_cuda and y.is_cuda):
        raise RuntimeError("ddm_b200.rbf_mmd2 runs on CUDA tensors only (no CPU fallback)")
    n, m = x.size(0), y.size(0)
    if n < 2 or m < 2:
        raise ValueError("Need at least two samples per set to compute MMD")
    if x.dim() != 2 or y.dim() != 2 or x.shape[1] != y.shape[1]:
        raise ValueError(f"expected x [n, D] and y [m, D], got {tuple(x.shape)} and {tuple(y.shape)}")
    if allow_tf32 is not None:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
        try:
            return rbf_mmd2(x, y, sigma, method="gemm")
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
    dtype = x.dtype
    xf, yf = x.detach().float().contiguous(), y.detach().float().contiguous()
    gamma = 1.0 / (2.0 * sigma**2)
    if method not in (None, "tc", "gemm"):
        raise ValueError("method must be None, 'tc' (fused tensor-core kernel) or 'gemm' (library Gram tiles + fused pass)")
    if method == "tc" or (method is None and allow_tf32 is None and min(n, m) * max(n, m) * x.shape[1] >= _TC_MIN_WORK):
        return _rbf_mmd2_tc(xf, yf, gamma).reshape(()).to(dtype)
    x2, y2 = ops.row_sqnorm(xf), ops.row_sqnorm(yf)
    kxx = _kernel_sum_offdiag(xf, x2, gamma) / (n * (n - 1))
    kyy = _kernel_sum_offdiag(yf, y2, gamma) / (m * (m - 1))
    kxy = _kernel_sum(xf, yf, x2, y2, gamma) / (n * m)
    return (kxx + kyy - 2.0 * kxy).reshape(()).to(dtype)
