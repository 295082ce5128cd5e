"""GPU parity of rbf_mmd2 (dddm/metrics.py:140-163): golden vectors from the reference, the fp64 oracle, and
size-independent properties at evaluation size."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(cuda_device):
    from ddm_b200 import _cabi

    _cabi.lib()
    return cuda_device


def test_rbf_mmd2_matches_reference_golden(dev, golden_mmd):
    import ddm_b200

    g = golden_mmd
    for n in [str(s) for s in g["names"]]:
        x, y, sigma = torch.from_numpy(g[f"{n}/x"]).to(dev), torch.from_numpy(g[f"{n}/y"]).to(dev), float(g[f"{n}/sigma"])
        got = ddm_b200.rbf_mmd2(x, y, sigma)
        assert got.dim() == 0 and got.dtype == torch.float32 and got.device == x.device
        ref, kxx, kyy, kxy = oracle.rbf_mmd2(g[f"{n}/x"], g[f"{n}/y"], sigma)
        scale = max(kxx, kyy, kxy, 1e-30)
        assert abs(float(got) - ref) <= 1e-5 * scale + 1e-30, (n, float(got), ref)
        assert abs(float(got) - float(g[f"{n}/mmd2_f64"])) <= 1e-5 * scale + 1e-30
    with pytest.raises(ValueError):
        ddm_b200.rbf_mmd2(torch.zeros(1, 2, device=dev), torch.zeros(3, 2, device=dev))
    with pytest.raises(RuntimeError):
        ddm_b200.rbf_mmd2(torch.zeros(4, 2), torch.zeros(3, 2))


def test_rbf_mmd2_kernel_sum_pieces(dev):
    """The fused pass against a dense fp64 evaluation: tiles with and without the diagonal, ragged widths, shifts."""
    from ddm_b200 import ops

    gen = torch.Generator().manual_seed(2)
    for rows, cols, D, shift, skip in ((37, 1029, 5, 0, True), (300, 300, 48, 0, True), (64, 5000, 16, 128, True),
                                      (513, 77, 3, 0, False), (1, 1, 2, 0, False)):
        a, b = torch.randn(rows, D, generator=gen), torch.randn(cols, D, generator=gen)
        gamma = 0.5 / D
        a2, b2 = ops.row_sqnorm(a.to(dev)), ops.row_sqnorm(b.to(dev))
        assert np.allclose(a2.cpu().numpy(), (a.double() ** 2).sum(-1).numpy(), rtol=2e-6)
        gram = a.to(dev) @ b.to(dev).t()
        got = float(ops.rbf_kernel_sum(gram, a2, b2, gamma, shift, skip))
        d2 = (a.double() ** 2).sum(-1)[:, None] + (b.double() ** 2).sum(-1)[None, :] - 2 * a.double() @ b.double().t()
        k = torch.exp(-gamma * d2)
        if skip:
            r = torch.arange(rows)[:, None] + shift
            k = torch.where(r == torch.arange(cols)[None, :], torch.zeros_like(k), k)
        assert abs(got - float(k.sum())) <= 2e-5 * float(k.sum()) + 1e-12, (rows, cols, got, float(k.sum()))


def test_rbf_mmd2_properties_at_evaluation_size(dev):
    """n = 9000 flattened images (two row chunks of the Gram matrix): symmetry, the closed form for identical sets,
    and agreement with a chunked fp64 evaluation on a subsample."""
    import ddm_b200

    gen = torch.Generator().manual_seed(4)
    x = (torch.rand(9000, 3072, generator=gen) * 2 - 1).to(dev)
    y = (torch.rand(8500, 3072, generator=gen) * 1.9 - 0.95).to(dev)
    sigma = 45.0
    a, b = ddm_b200.rbf_mmd2(x, y, sigma), ddm_b200.rbf_mmd2(y, x, sigma)
    assert abs(float(a) - float(b)) <= 1e-6
    same = float(ddm_b200.rbf_mmd2(x, x.clone(), sigma))
    # identical sets: kxy's mean includes the n diagonal ones, kxx's does not: mmd2 = 2 (kxx - kxy), kxy = ((n-1) kxx + 1) / n
    n = x.shape[0]
    sub = x[:700].double().cpu().numpy()
    _, kxx_sub, _, _ = oracle.rbf_mmd2(sub, sub, sigma)
    assert same < 0 and abs(same - 2.0 * (kxx_sub - ((n - 1) * kxx_sub + 1.0) / n)) <= 5e-3 * abs(same)
    ref, kxx, kyy, kxy = oracle.rbf_mmd2(x[:600].cpu().numpy(), y[:500].cpu().numpy(), sigma)
    got = float(ddm_b200.rbf_mmd2(x[:600], y[:500], sigma))
    assert abs(got - ref) <= 1e-5 * max(kxx, kyy, kxy)


def test_rbf_kernel_sum_tensor_core_pieces(dev):
    """The fused tcgen05 kernel (bf16 hi/lo split, accumulator read from tensor memory) against a dense fp64 evaluation:
    tiles that end inside a 128 x 256 block, D that is not a multiple of 64, one tile, many tiles per CTA, the symmetric
    (strict upper triangle counted twice) and the rectangular form; deterministic."""
    from ddm_b200 import ops

    gen = torch.Generator().manual_seed(7)
    for rows, cols, D, sym in ((128, 256, 64, False), (300, 300, 130, True), (37, 1029, 5, False), (1000, 777, 3072, False),
                               (513, 513, 200, True), (2, 2, 2, True), (129, 257, 65, False), (4000, 4000, 96, True)):
        a = torch.randn(rows, D, generator=gen)
        b = a if sym else torch.randn(cols, D, generator=gen)
        gamma = 0.5 / D
        ad, bd = a.to(dev), b.to(dev)
        a2, b2 = ops.row_sqnorm(ad), ops.row_sqnorm(bd)
        ah, al = ops.rbf_split_bf16(ad)
        bh, bl = (ah, al) if sym else ops.rbf_split_bf16(bd)
        assert ah.shape == (rows, (D + 63) // 64 * 64) and float((ah[:, D:].float().abs().sum())) == 0.0
        rec = ah[:, :D].float() + al[:, :D].float()
        assert float((rec - ad).abs().max()) <= 2.0 ** -16 * float(ad.abs().max())
        got = float(ops.rbf_kernel_sum_tc(ah, al, bh, bl, a2, b2, D, gamma, sym))
        again = float(ops.rbf_kernel_sum_tc(ah, al, bh, bl, a2, b2, D, gamma, sym))
        assert got == again
        d2 = (a.double() ** 2).sum(-1)[:, None] + (b.double() ** 2).sum(-1)[None, :] - 2 * a.double() @ b.double().t()
        k = torch.exp(-gamma * d2)
        if sym:
            k = k - torch.diag(torch.diag(k))
        assert abs(got - float(k.sum())) <= 2e-5 * float(k.sum()) + 1e-12, (rows, cols, D, sym, got, float(k.sum()))


def test_rbf_mmd2_fused_path_equals_gemm_path(dev, golden_mmd):
    """method='tc' against method='gemm' (fp32 library Gram) and the fp64 oracle at 2000 x 1500 x 3072; the reference's
    golden cases also pass through the fused kernel within the budget of its bf16x2 split."""
    import ddm_b200

    gen = torch.Generator().manual_seed(9)
    x = (torch.rand(2000, 3072, generator=gen) * 2 - 1).to(dev)
    y = (torch.rand(1500, 3072, generator=gen) * 1.9 - 0.95).to(dev)
    tc, gm = float(ddm_b200.rbf_mmd2(x, y, 45.0, method="tc")), float(ddm_b200.rbf_mmd2(x, y, 45.0, method="gemm"))
    ref, kxx, kyy, kxy = oracle.rbf_mmd2(x[:400].cpu().numpy(), y[:300].cpu().numpy(), 45.0)
    assert abs(tc - gm) <= 2e-6 * max(kxx, kyy, kxy), (tc, gm)
    sub = float(ddm_b200.rbf_mmd2(x[:400], y[:300], 45.0, method="tc"))
    assert abs(sub - ref) <= 1e-5 * max(kxx, kyy, kxy), (sub, ref)
    assert float(ddm_b200.rbf_mmd2(x, y, 45.0)) == tc  # evaluation-sized sets take the fused path by default
    g = golden_mmd
    for n in [str(s) for s in g["names"]]:
        xs, ys, sigma = torch.from_numpy(g[f"{n}/x"]).to(dev), torch.from_numpy(g[f"{n}/y"]).to(dev), float(g[f"{n}/sigma"])
        ref, kxx, kyy, kxy = oracle.rbf_mmd2(g[f"{n}/x"], g[f"{n}/y"], sigma)
        got = float(ddm_b200.rbf_mmd2(xs, ys, sigma, method="tc"))
        assert abs(got - ref) <= 2e-4 * max(kxx, kyy, kxy, 1e-30), (n, got, ref)
