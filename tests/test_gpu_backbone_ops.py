"""GPU parity of the backbone helper kernels (csrc/backbone_ops.cu): LayerNorm fwd/bwd and column sums against
plain PyTorch fp32 references of the same op, and the DiT with the kernels on vs. the ATen composition."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(cuda_device):
    from ddm_b200 import _cabi

    _cabi.lib()
    return cuda_device


@pytest.mark.parametrize("N,C", [(1, 4), (7, 384), (1000, 384), (333, 1024), (65, 96), (4096, 128), (5, 1000)])
def test_layer_norm_matches_torch(dev, N, C):
    from ddm_b200 import ops

    gen = torch.Generator().manual_seed(N + C)
    x = (torch.randn(N, C, generator=gen) * 1.7 + 0.3).to(dev)
    w = (1.0 + 0.2 * torch.randn(C, generator=gen)).to(dev)
    b = (0.1 * torch.randn(C, generator=gen)).to(dev)
    gy = torch.randn(N, C, generator=gen).to(dev)
    for dtype, tol in ((torch.float32, 2e-6), (torch.bfloat16, 1.2e-2)):
        xr, wr, br = (t.detach().to(dtype).double().requires_grad_(True) for t in (x, w, b))  # fp64 reference, rounded inputs
        ref = F.layer_norm(xr, (C,), wr, br, 1e-5)
        ref.backward(gy.to(dtype).double())
        xa, wa, ba = (t.detach().clone().to(dtype).requires_grad_(True) for t in (x, w, b))
        if dtype == torch.bfloat16 and (C * 2) % 16:
            assert not ops.layer_norm_supported(xa, wa, ba)  # rows must be 16-byte aligned
            continue
        assert ops.layer_norm_supported(xa, wa, ba)
        y, mean, rstd = ops.layer_norm(xa, wa, ba, 1e-5)
        y.backward(gy.to(dtype))
        assert y.dtype == dtype and mean.dtype == torch.float32

        def rel(a, r):
            return float((a.double() - r).abs().max() / r.abs().max().clamp_min(1e-12))

        assert rel(y, ref) <= tol, (dtype, rel(y, ref))
        assert rel(mean, xr.detach().mean(-1)) <= 1e-5
        assert rel(xa.grad, xr.grad) <= 2 * tol, (dtype, "dx", rel(xa.grad, xr.grad))
        wtol = tol if dtype == torch.float32 else 2e-2
        assert rel(wa.grad, wr.grad) <= 3 * wtol, (dtype, "dw", rel(wa.grad, wr.grad))
        assert rel(ba.grad, br.grad) <= 3 * wtol, (dtype, "db", rel(ba.grad, br.grad))
        y2, _, _ = ops.layer_norm(xa, wa, ba, 1e-5)  # deterministic
        assert torch.equal(y2, y)
    assert not ops.layer_norm_supported(torch.zeros(3, 6, device=dev), torch.zeros(6, device=dev), torch.zeros(6, device=dev))


@pytest.mark.parametrize("N,C", [(1, 4), (9, 384), (70000, 384), (1025, 1536), (300, 48), (4099, 1152)])
def test_colsum_matches_torch(dev, N, C):
    from ddm_b200 import ops

    gen = torch.Generator().manual_seed(N * 3 + C)
    a = torch.randn(N, C, generator=gen).to(dev)
    ref = a.double().sum(0)
    out = ops.colsum(a)
    scale = float(a.double().abs().sum(0).max())
    assert float((out.double() - ref).abs().max()) <= 1e-6 * scale
    ab = a.bfloat16()
    if (C * 2) % 16:
        assert not ops.colsum_supported(ab)
        return
    outb = ops.colsum(ab)
    refb = ab.double().sum(0)
    assert outb.dtype == torch.bfloat16
    assert float((outb.double() - refb).abs().max()) <= 8e-3 * float(refb.abs().max()) + 1e-6 * scale
    assert torch.equal(ops.colsum(a), out)  # fixed summation order


def test_dit_kernels_on_equals_off(dev):
    """The DiT with the LayerNorm / column-sum kernels gives the same output and gradients as the ATen composition."""
    from ddm_b200 import backbones

    for dtype, tol in ((torch.float32, 2e-4), (torch.bfloat16, 4e-2)):
        torch.manual_seed(0)
        model = backbones.DDDMDiT(depth=2, embed_dim=96, num_heads=3).to(dev).to(dtype)
        xt, xi, t = torch.randn(12, 3, 32, 32, device=dev), torch.randn(12, 3, 32, 32, device=dev), torch.rand(12, device=dev)
        res = {}
        try:
            for on in (True, False):
                backbones.USE_CUDA_KERNELS = on
                model.zero_grad(set_to_none=True)
                y = model(xt, t, xi)
                (y.float() ** 2).mean().backward()
                res[on] = (y.detach().float(), torch.cat([p.grad.float().reshape(-1) for p in model.parameters()]))
        finally:
            backbones.USE_CUDA_KERNELS = True
        (ya, ga), (yb, gb) = res[True], res[False]
        assert float((ya - yb).abs().max()) <= tol * float(yb.abs().max())
        assert float((ga - gb).abs().max()) <= tol * float(gb.abs().max()), (dtype, float((ga - gb).abs().max()), float(gb.abs().max()))
