"""GPU parity of the elementwise kernels (K2, K3, K4), the training step and the sampler."""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from tests.helpers import MixModel

pytestmark = pytest.mark.gpu


def _names(g):
    return [str(n) for n in g["names"]]


@pytest.fixture(scope="module")
def dev(cuda_device):
    from ddm_b200 import _cabi

    _cabi.lib()
    return cuda_device


def test_sigmoid_weight(dev, golden_weights):
    from ddm_b200 import ops, sigmoid_weight

    g = golden_weights
    t = torch.from_numpy(g["t"]).float().to(dev)
    for bias in (0.0, 1.0, -0.5):
        w = sigmoid_weight(t, bias)
        assert w.shape == t.shape and w.dtype == t.dtype
        assert np.allclose(w.cpu().numpy(), g[f"w32_bias{bias}"], rtol=2e-6, atol=1e-12)
        assert np.allclose(w.cpu().numpy(), g[f"w64_bias{bias}"], rtol=2e-6, atol=1e-12)
        _, s = ops.sigmoid_weight_sum(t, bias)
        assert abs(float(s) - g[f"w64_bias{bias}"].sum()) <= 2e-6 * g[f"w64_bias{bias}"].sum()
    big = torch.rand(5000, device=dev)
    w, s = ops.sigmoid_weight_sum(big, 0.3)
    assert abs(float(s) - oracle.sigmoid_weight(big.cpu().numpy(), 0.3).sum()) <= 1e-5 * 5000
    a = ops.sigmoid_weight_sum(big, 0.3)[1]
    assert torch.equal(a, s)  # fixed-shape reduction: bitwise reproducible


def test_forward_marginal_bit_exact(dev, golden_schedules):
    from ddm_b200 import forward_marginal_sample, ops

    g = golden_schedules
    for name in ("img", "flat", "toy", "lowrank"):
        x0, eps, t = (torch.from_numpy(g[f"fm_{name}/{k}"]).to(dev) for k in ("x0", "eps", "t"))
        xt = forward_marginal_sample(x0, t, eps)
        assert xt.shape == x0.shape
        assert np.array_equal(xt.cpu().numpy(), g[f"fm_{name}/xt"]), name  # same fp32 ops as the reference's eager path
    x0, eps, t = (torch.from_numpy(g[f"fm_img/{k}"]).to(dev) for k in ("x0", "eps", "t"))
    xt, rep = ops.forward_marginal_expand(x0, t, eps, 5, True)
    assert rep.shape == (4 * 5, 3, 2, 2)
    assert torch.equal(rep.view(4, 5, 3, 2, 2), xt[:, None].expand(4, 5, 3, 2, 2))
    # CIFAR-sized, vectorised path, fp32 and bf16, against the oracle
    gen = torch.Generator().manual_seed(0)
    x0 = (torch.rand(128, 3, 32, 32, generator=gen) * 2 - 1).to(dev)
    eps = torch.randn(128, 3, 32, 32, generator=gen).to(dev)
    t = torch.rand(128, generator=gen).to(dev)
    _, rep = ops.forward_marginal_expand(x0, t, eps, 8, False)
    ref, ref_rep = oracle.forward_marginal(x0, t, eps, m=8)
    assert np.allclose(rep.cpu().numpy().reshape(1024, -1), ref_rep, rtol=0, atol=3e-7)
    _, rep16 = ops.forward_marginal_expand(x0.bfloat16(), t, eps.bfloat16(), 8, False)
    ref16, ref16_rep = oracle.forward_marginal(x0.bfloat16(), t, eps.bfloat16(), m=8)
    assert np.allclose(rep16.float().cpu().numpy().reshape(1024, -1), ref16_rep, rtol=8e-3, atol=1e-3)


def test_forward_marginal_autograd(dev):
    from ddm_b200 import forward_marginal_sample

    x0 = torch.randn(4, 6, device=dev, requires_grad=True)
    eps = torch.randn(4, 6, device=dev, requires_grad=True)
    t = torch.rand(4, device=dev, requires_grad=True)
    y = forward_marginal_sample(x0, t, eps)
    gy = torch.randn_like(y)
    y.backward(gy)
    tt = t.detach()[:, None]
    assert torch.allclose(x0.grad, (1 - tt) * gy) and torch.allclose(eps.grad, tt * gy)
    assert torch.allclose(t.grad, ((eps.detach() - x0.detach()) * gy).sum(1), atol=1e-6)


def test_bridge_bit_exact(dev, golden_schedules):
    from ddm_b200 import gaussian_bridge_mu_sigma, ops

    g = golden_schedules
    x0hat, xt = torch.from_numpy(g["br/x0hat"]).to(dev), torch.from_numpy(g["br/xt"]).to(dev)
    for steps in (20, 5):
        grid = torch.linspace(0.0, 1.0, steps + 1, device=dev)
        for churn in (1.0, 0.0, 0.5):
            key = f"br_grid{steps}_churn{churn}"
            for k in range(steps):
                mu, std = gaussian_bridge_mu_sigma(grid[k], grid[k + 1], x0hat, xt, eps_churn=churn)
                assert std.shape == (1, 1)
                assert np.array_equal(mu.cpu().numpy(), g[f"{key}/mu"][k]), (key, k)
                assert float(std) == g[f"{key}/std"][k], (key, k)
    s, t = torch.from_numpy(g["br_vec/s"]).to(dev), torch.from_numpy(g["br_vec/t"]).to(dev)
    for churn in (1.0, 0.0, 0.7):
        mu, std = gaussian_bridge_mu_sigma(s, t, x0hat, xt, eps_churn=churn)
        assert std.shape == (5, 1)
        assert np.array_equal(mu.cpu().numpy(), g[f"br_vec_churn{churn}/mu"])
        assert np.array_equal(std.cpu().numpy(), g[f"br_vec_churn{churn}/std"])
        z = torch.randn_like(xt)
        nxt = ops.bridge_step(xt, x0hat, z, s, t, churn)
        assert torch.equal(nxt, mu + std * z)
    mu, std = gaussian_bridge_mu_sigma(s, t, torch.from_numpy(g["br_img/x0hat"]).to(dev),
                                       torch.from_numpy(g["br_img/xt"]).to(dev), eps_churn=1.0)
    assert std.shape == (5, 1, 1, 1) and np.array_equal(mu.cpu().numpy(), g["br_img/mu"])
    assert np.array_equal(std.cpu().numpy(), g["br_img/std"])
    # image-sized vectorised path against the fp64 oracle, fp32 and bf16
    x = torch.randn(128, 3, 32, 32, device=dev)
    h = torch.randn_like(x)
    z = torch.randn_like(x)
    sc, tc = torch.tensor([0.45], device=dev), torch.tensor([0.5], device=dev)
    out = ops.bridge_step(x, h, z, sc, tc, 0.8)
    ref, _ = oracle.bridge_step(x, h, z, float(sc), float(tc), 0.8)
    assert np.allclose(out.cpu().numpy(), ref, rtol=0, atol=2e-6)
    out16 = ops.bridge_step(x.bfloat16(), h.bfloat16(), z.bfloat16(), sc, tc, 0.8)
    ref16, _ = oracle.bridge_step(x.bfloat16(), h.bfloat16(), z.bfloat16(), float(sc), float(tc), 0.8)
    assert np.allclose(out16.float().cpu().numpy(), ref16, rtol=8e-3, atol=2e-3)


def test_training_step_matches_reference(dev, golden_step):
    """The drop-in step with the recorded noise passed in reproduces the reference's loss, metrics
    and gradients (w.r.t. the denoiser output and the model parameters)."""
    from ddm_b200 import distributional_training_step

    g = golden_step
    for n in _names(g):
        m, beta, lam, w_bias = g[f"{n}/hyper"]
        m = int(m)
        model = MixModel().to(dev)
        x0, t, eps, xi = (torch.from_numpy(g[f"{n}/{k}"]).to(dev) for k in ("x0", "t", "eps", "xi"))
        seen = {}

        def hook(_mod, args, out):
            seen["xt_rep"], seen["t_rep"], seen["xi"] = (a.detach() for a in args)
            out.retain_grad()
            seen["out"] = out

        h = model.register_forward_hook(hook)
        loss, metrics = distributional_training_step(model, x0, m=m, beta=float(beta), lam=float(lam),
                                                     w_bias=float(w_bias), t=t, eps=eps, xi=xi)
        h.remove()
        loss.backward()
        assert loss.dim() == 0 and loss.dtype == x0.dtype
        assert list(metrics) == ["loss", "confidence", "interaction", "weight"]
        assert all(isinstance(v, float) for v in metrics.values())
        got = np.array([metrics[k] for k in ("loss", "confidence", "interaction", "weight")])
        assert np.allclose(got, g[f"{n}/scalars"], rtol=1e-5, atol=1e-7), (n, got, g[f"{n}/scalars"])
        B = x0.shape[0]
        assert np.array_equal(seen["xt_rep"].view(B, m, *x0.shape[1:])[:, 0].cpu().numpy(), g[f"{n}/xt"])
        assert np.allclose(seen["out"].detach().cpu().numpy().reshape(g[f"{n}/xhat"].shape), g[f"{n}/xhat"],
                           rtol=1e-5, atol=1e-6)
        gref = g[f"{n}/grad_xhat"]
        gx = seen["out"].grad.cpu().numpy().reshape(gref.shape)
        assert np.max(np.abs(gx - gref)) <= 2e-5 * np.max(np.abs(gref)), n
        pg = np.array([float(model.a.grad), float(model.b.grad), float(model.c.grad)])
        assert np.allclose(pg, g[f"{n}/param_grads"], rtol=2e-4, atol=1e-7), (n, pg, g[f"{n}/param_grads"])
    with pytest.raises(ValueError, match="m must be >= 2"):
        distributional_training_step(MixModel().to(dev), torch.zeros(2, 2, device=dev), m=1, beta=1.0, lam=1.0,
                                     w_bias=0.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        distributional_training_step(MixModel(), torch.zeros(2, 2), m=2, beta=1.0, lam=1.0, w_bias=0.0)


def test_training_step_rng_order_and_deferred_metrics(dev):
    """Seeded run draws rand(B), randn_like(x0), randn(B,m,...) in the reference's order (training.py:65-69)."""
    from ddm_b200 import DeferredMetrics, distributional_training_step

    model = MixModel().to(dev)
    x0 = torch.randn(8, 3, 4, 4, device=dev).clamp(-1, 1)
    torch.manual_seed(7)
    loss_a, met = distributional_training_step(model, x0, m=4, beta=0.1, lam=1.0, w_bias=0.0, sync_metrics=False)
    assert isinstance(met, DeferredMetrics) and met.tensor.is_cuda
    torch.manual_seed(7)
    t = torch.rand(8, device=dev)
    eps = torch.randn_like(x0)
    xi = torch.randn((8, 4, 3, 4, 4), device=dev)
    loss_b, met_b = distributional_training_step(model, x0, m=4, beta=0.1, lam=1.0, w_bias=0.0, t=t, eps=eps, xi=xi)
    assert torch.equal(loss_a, loss_b)
    assert met["loss"] == met_b["loss"] and dict(met) == met_b
    # x0.requires_grad routes through the split kernels and still matches
    x0g = x0.clone().requires_grad_(True)
    loss_c, _ = distributional_training_step(model, x0g, m=4, beta=0.1, lam=1.0, w_bias=0.0, t=t, eps=eps, xi=xi)
    loss_c.backward()
    assert abs(float(loss_c) - float(loss_b)) <= 1e-5 * abs(float(loss_b)) and x0g.grad is not None


def test_sampler_matches_reference(dev, golden_sampler):
    from ddm_b200 import sample_dddm

    g = golden_sampler
    for n in _names(g):
        steps, churn = g[f"{n}/hyper"]
        model = MixModel().train()
        noise = tuple(torch.from_numpy(g[f"{n}/{k}"]) for k in ("x_init", "xis", "zs"))
        x = sample_dddm(model, n_samples=noise[0].shape[0], steps=int(steps), eps_churn=float(churn), device=str(dev),
                        data_shape=noise[0].shape[1:], noise=noise)
        assert not model.training and next(model.parameters()).is_cuda
        assert x.shape == g[f"{n}/x_final"].shape
        # the bridge is bit-exact; MixModel's tanh differs by ulps between CPU and GPU libm
        assert np.allclose(x.cpu().numpy(), g[f"{n}/x_final"], rtol=2e-5, atol=2e-5), n
    torch.manual_seed(3)
    a = sample_dddm(MixModel(), n_samples=16, steps=3, device=str(dev))
    torch.manual_seed(3)
    x = torch.randn(16, 2, device=dev)
    assert a.shape == (16, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        sample_dddm(MixModel(), n_samples=2, steps=1, device="cpu")


def test_host_session_matches_device_path(dev):
    """C-ABI with HOST buffers (bench.py's e2e path) == device path == oracle."""
    from ddm_b200 import _cabi

    L = _cabi.lib()
    B, m, D = 32, 8, 768
    gen = torch.Generator().manual_seed(2)
    x0 = torch.randn(B, D, generator=gen).clamp(-1, 1)
    xh = x0[:, None] + 0.05 * torch.randn(B, m, D, generator=gen)
    t = torch.rand(B, generator=gen)
    w = oracle.sigmoid_weight(t.numpy(), 0.2).mean()
    loss, conf, inter, grad = oracle.energy_loss(xh.numpy(), x0.numpy(), 0.1, 1.0, w)
    for dtype, code, rel in ((torch.float32, 0, 1e-5), (torch.bfloat16, 1, 1e-2)):
        a, c = xh.to(dtype).contiguous().pin_memory(), x0.to(dtype).contiguous().pin_memory()
        if dtype == torch.bfloat16:
            loss, conf, inter, grad = oracle.energy_loss(a, c, 0.1, 1.0, w)
        gout = torch.empty_like(a).pin_memory()
        out = torch.empty(4, dtype=torch.float32).pin_memory()
        s = L.dddm_session_create(B, m, D, code, dev.index or 0)
        assert s
        try:
            for _ in range(4):  # more steps than slots: exercises buffer rotation
                _cabi.check(L.dddm_session_step_host(s, a.data_ptr(), c.data_ptr(), t.data_ptr(), 0.2, 0.1, 1.0,
                                                     gout.data_ptr(), out.data_ptr()))
            got = out.numpy().astype(np.float64)
            assert np.allclose(got, [loss, conf, inter, w], rtol=2e-5, atol=1e-7)
            assert np.max(np.abs(gout.float().numpy() - grad)) <= rel * np.max(np.abs(grad))
            outs = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(7)]
            for o in outs:
                _cabi.check(L.dddm_session_enqueue_host(s, a.data_ptr(), c.data_ptr(), t.data_ptr(), 0.2, 0.1, 1.0,
                                                        None, o.data_ptr()))
            _cabi.check(L.dddm_session_wait(s))
            assert all(torch.equal(o, out) for o in outs)
            # packed host layout: [xhat | x0 | t] and [grad | out] in ONE pinned buffer each -> one copy per direction
            sz = [ctypes.c_size_t() for _ in range(5)]
            _cabi.check(L.dddm_session_packed_layout(s, *[ctypes.addressof(v) for v in sz]))
            in_bytes, x0_off, t_off, out_bytes, out_off = [int(v.value) for v in sz]
            assert x0_off >= a.numel() * a.element_size() and t_off >= x0_off + c.numel() * c.element_size()
            assert x0_off % 256 == 0 and t_off % 256 == 0 and out_off % 256 == 0 and in_bytes >= t_off + 4 * B
            pin, pout = torch.zeros(in_bytes, dtype=torch.uint8).pin_memory(), torch.zeros(out_bytes, dtype=torch.uint8).pin_memory()
            pin[:a.numel() * a.element_size()] = a.view(torch.uint8).reshape(-1)
            pin[x0_off:x0_off + c.numel() * c.element_size()] = c.view(torch.uint8).reshape(-1)
            pin[t_off:t_off + 4 * B] = t.view(torch.uint8).reshape(-1)
            for _ in range(6):
                _cabi.check(L.dddm_session_enqueue_host(s, pin.data_ptr(), pin.data_ptr() + x0_off, pin.data_ptr() + t_off,
                                                        0.2, 0.1, 1.0, pout.data_ptr(), pout.data_ptr() + out_off))
            _cabi.check(L.dddm_session_wait(s))
            assert torch.equal(pout[out_off:out_off + 16].view(torch.float32), out)
            assert torch.equal(pout[:a.numel() * a.element_size()].view(dtype).reshape(a.shape), gout)
        finally:
            L.dddm_session_destroy(s)
    assert not L.dddm_session_create(0, 8, 4, 0, 0) and L.dddm_last_error() == -2
    for alloc in (L.dddm_host_alloc, L.dddm_host_alloc_input):  # pinned / pinned write-combined (inputs only)
        p = alloc(1024)
        assert p
        ctypes.memset(p, 0, 1024)
        L.dddm_host_free(p)


def test_launch_counter(dev):
    from ddm_b200 import _cabi, ops

    before = _cabi.launch_count()
    ops.sigmoid_weight_sum(torch.rand(16, device=dev), 0.0)
    assert _cabi.launch_count() == before + 1


def test_sharded_ranks_equal_global_batch(dev):
    """Emulate R data-parallel ranks on one GPU by looping shards: with the global
    w-sum (what the all-reduce produces) the rank-mean of per-shard K1 gradients is the global-batch gradient."""
    from ddm_b200 import ops

    B, m, D, R = 32, 8, 3072, 4
    gen = torch.Generator().manual_seed(9)
    x0 = torch.randn(B, D, generator=gen).clamp(-1, 1).to(dev)
    xh = (x0[:, None].cpu() + 0.05 * torch.randn(B, m, D, generator=gen)).to(dev)
    t = torch.rand(B, generator=gen).to(dev)
    _, w_sum = ops.sigmoid_weight_sum(t, 0.0)
    out_g, grad_g = ops.energy_fused(xh, x0, w_sum, 1.0 / B, 0.1, 1.0, True)
    per = B // R
    shard_sums = [ops.sigmoid_weight_sum(t[r * per:(r + 1) * per], 0.0)[1] for r in range(R)]
    reduced = torch.stack(shard_sums).sum(dim=0)  # == all_reduce(SUM)
    assert abs(float(reduced) - float(w_sum)) <= 1e-6 * float(w_sum)
    grads, losses = [], []
    for r in range(R):
        sl = slice(r * per, (r + 1) * per)
        o, g = ops.energy_fused(xh[sl].contiguous(), x0[sl].contiguous(), reduced, 1.0 / (R * per), 0.1, 1.0, True)
        grads.append(g / R)  # DDP averages over ranks
        losses.append(o[0])
    got = torch.cat(grads, dim=0)
    assert float((got - grad_g).abs().max()) <= 2e-6 * float(grad_g.abs().max())
    assert abs(float(torch.stack(losses).mean()) - float(out_g[0])) <= 2e-6 * abs(float(out_g[0]))


def _patchify(x, p):
    """[B,C,H,W] -> [B, (H/p)(W/p), C*p*p]: the layout of PatchUnembed's projection (dddm/model.py:125-128)."""
    B, C, H, W = x.shape
    return x.view(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), C * p * p)


def test_forward_marginal_concat_bit_exact(dev):
    """K2c == cat(forward_marginal_sample(...) expanded m-fold, xi) bit for bit (fp32), RN-rounded for a bf16
    backbone, and x0 in patch-token order (dddm/training.py:66-73, dddm/model.py:236)."""
    from ddm_b200 import ops
    from oracle import torch_port

    gen = torch.Generator().manual_seed(3)
    # m = 8: one unguarded block of 8 draws; m = 11: that block + a guarded remainder; m = 2, 3: the guarded block alone
    for B, m, C, H, W, p in ((5, 3, 3, 32, 32, 4), (2, 8, 3, 16, 24, 8), (3, 2, 1, 8, 8, 4), (2, 11, 3, 8, 8, 4)):
        x0 = (torch.rand(B, C, H, W, generator=gen) * 2 - 1)
        eps, xi, t = torch.randn(B, C, H, W, generator=gen), torch.randn(B, m, C, H, W, generator=gen), torch.rand(B, generator=gen)
        xt = torch_port.forward_marginal(x0, t, eps)  # CPU eager fp32 = the reference's arithmetic
        want = torch.cat([xt[:, None].expand(B, m, C, H, W).reshape(B * m, C, H, W), xi.reshape(B * m, C, H, W)], dim=1)
        x6, tok = ops.forward_marginal_concat(x0.to(dev), t.to(dev), eps.to(dev), xi.to(dev), False, p)
        assert x6.dtype == torch.float32 and torch.equal(x6.cpu(), want)
        assert torch.equal(tok.cpu().view(B, -1), _patchify(x0, p).reshape(B, -1))
        x6b, tok0 = ops.forward_marginal_concat(x0.to(dev), t.to(dev), eps.to(dev), xi.to(dev), True, 0)
        assert x6b.dtype == torch.bfloat16 and tok0.numel() == 0 and torch.equal(x6b.cpu(), want.bfloat16())
        xb = [v.bfloat16() for v in (x0, eps, xi)]
        x6c, tokc = ops.forward_marginal_concat(xb[0].to(dev), t.to(dev), xb[1].to(dev), xb[2].to(dev), False, p)
        xtb = ((1 - t).view(B, 1, 1, 1) * xb[0].float() + t.view(B, 1, 1, 1) * xb[1].float())
        wantb = torch.cat([xtb[:, None].expand(B, m, C, H, W).reshape(B * m, C, H, W),
                           xb[2].float().reshape(B * m, C, H, W)], dim=1).bfloat16()
        assert x6c.dtype == torch.bfloat16 and torch.equal(x6c.cpu(), wantb)
        assert torch.equal(tokc.cpu().view(B, -1), _patchify(xb[0], p).reshape(B, -1))
    with pytest.raises(ValueError):
        ops.forward_marginal_concat(x0.to(dev), t.to(dev), eps.to(dev), xi[:, :, :, :4].to(dev), False, 4)


def test_fused_io_step_equals_generic_step(dev):
    """The DiT fast path (K2c -> forward_cat(tokens=True) -> K1 on patch-token order) gives the same loss, metrics
    and parameter gradients as the generic path (K2 -> model(x_t, t, xi) -> K1), in fp32 and with a bf16 backbone."""
    import ddm_b200
    from ddm_b200.backbones import DDDMDiT

    torch.manual_seed(0)
    B, m = 6, 4
    x0 = (torch.rand(B, 3, 32, 32) * 2 - 1).to(dev)
    t, eps, xi = torch.rand(B).to(dev), torch.randn(B, 3, 32, 32).to(dev), torch.randn(B, m, 3, 32, 32).to(dev)
    for dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 3e-2)):
        torch.manual_seed(1)
        model = DDDMDiT(depth=2, embed_dim=96, num_heads=3).to(dev).to(dtype)
        torch.nn.init.normal_(model.unembed.proj.weight, std=0.05)
        res = {}
        for fused in (True, False):
            model.zero_grad(set_to_none=True)
            loss, met = ddm_b200.distributional_training_step(model, x0, m=m, beta=0.1, lam=1.0, w_bias=0.0, t=t, eps=eps,
                                                              xi=xi, fused_io=fused)
            loss.backward()
            res[fused] = (float(loss), met, torch.cat([p.grad.float().reshape(-1) for p in model.parameters()]))
        (la, ma, ga), (lb, mb, gb) = res[True], res[False]
        assert abs(la - lb) <= tol * abs(lb), (dtype, la, lb)
        assert all(abs(ma[k] - mb[k]) <= tol * max(abs(mb[k]), 1e-6) for k in ma)
        assert float((ga - gb).abs().max()) <= tol * float(gb.abs().max()), (dtype, float((ga - gb).abs().max()))
    with pytest.raises(ValueError):
        ddm_b200.distributional_training_step(MixModel().to(dev), x0, m=m, beta=0.1, lam=1.0, w_bias=0.0, fused_io=True)


def test_bridge_step_philox_matches_torch_randn(dev):
    """K3 with the draws fused in: z (in registers) and the next xi (written out) are bit-identical to
    torch.randn_like at the same generator (seed, offset), for sizes below / at / above ATen's grid cap, ragged
    sizes, fp32 and bf16, offsets by value and from device memory; the update equals K3 fed with torch's z."""
    from ddm_b200 import ops

    gen = torch.cuda.default_generators[dev.index or 0]
    for dtype in (torch.float32, torch.bfloat16):
        for shape in ((5, 7), (64, 3, 8, 8), (128, 3072), (1024, 3, 32, 32), (1000, 1213)):
            torch.manual_seed(2024 + len(shape))
            x = torch.randn(shape, device=dev).to(dtype)
            xh = torch.randn(shape, device=dev).to(dtype)
            s, t = torch.tensor([0.35], device=dev), torch.tensor([0.4], device=dev)
            numel = x.numel()
            c = ops.philox_increment(numel, dev)
            off = gen.get_offset()
            z_ref = torch.randn_like(x)
            assert gen.get_offset() == off + c, (shape, gen.get_offset() - off, c)  # ATen advances by what we predict
            xi_ref = torch.randn_like(x)
            want = ops.bridge_step(x, xh, z_ref, s, t, 0.8)
            for via_device in (False, True):
                xa, xi = x.clone(), torch.empty_like(x)
                if via_device:
                    ph = torch.tensor([gen.initial_seed(), off, off + c], dtype=torch.int64, device=dev)
                    ops.bridge_step_philox_(xa, xh, xi, s, t, 0.8, philox=ph)
                else:
                    ops.bridge_step_philox_(xa, xh, xi, s, t, 0.8, seed=gen.initial_seed(), offset_z=off, offset_xi=off + c)
                assert torch.equal(xi, xi_ref), (dtype, shape, via_device, float((xi.float() - xi_ref.float()).abs().max()))
                assert torch.equal(xa, want), (dtype, shape, via_device, float((xa.float() - want.float()).abs().max()))
            xb = x.clone()  # last step of a run: no next xi
            ops.bridge_step_philox_(xb, xh, None, s, t, 0.8, seed=gen.initial_seed(), offset_z=off)
            assert torch.equal(xb, want)
    with pytest.raises(ValueError):
        ops.bridge_step_philox_(x, xh, None, torch.rand(1000, device=dev), torch.rand(1000, device=dev), 0.8)


def test_sampler_fused_noise_equals_reference_draw_order(dev):
    """sample_dddm with the draws fused into K3 (eager and graphed) == the same run with every draw made by
    torch.randn in the reference's order (x_T, then per step xi, z; dddm/sampling.py:23-31) and passed in;
    the generator ends at the same offset."""
    import ddm_b200

    model = MixModel().to(dev)
    for shape, n, steps in (((3, 8, 8), 64, 7), ((3, 32, 32), 512, 4), ((2,), 33, 5)):
        torch.manual_seed(77)
        xT = torch.randn((n, *shape), device=dev)
        xis, zs = [None] * steps, [None] * steps
        for k in reversed(range(steps)):
            xis[k] = torch.randn_like(xT)
            zs[k] = torch.randn_like(xT)
        tail_ref = torch.randn(4, device=dev)
        ref = ddm_b200.sample_dddm(model, n_samples=n, steps=steps, eps_churn=0.7, device=str(dev), data_shape=shape,
                                   noise=(xT, xis, zs))
        for graph in (False, True):
            for fused in (True, False):
                torch.manual_seed(77)
                out = ddm_b200.sample_dddm(model, n_samples=n, steps=steps, eps_churn=0.7, device=str(dev), data_shape=shape,
                                           cuda_graph=graph, fused_noise=fused)
                tail = torch.randn(4, device=dev)
                assert torch.equal(out, ref), (shape, graph, fused, float((out - ref).abs().max()))
                assert torch.equal(tail, tail_ref), (shape, graph, fused)


def test_nvtx_ranges_toggle(dev):
    """NVTX ranges around the C-ABI entry points (tracing) can be switched on and off at run time; with no
    profiler attached they are no-ops and results do not change."""
    from ddm_b200 import _cabi, ops

    t = torch.rand(64, device=dev)
    base = ops.sigmoid_weight_sum(t, 0.0)[1].clone()
    try:
        _cabi.set_tuning("nvtx", 1)
        assert _cabi.get_tuning("nvtx") == 1
        assert torch.equal(ops.sigmoid_weight_sum(t, 0.0)[1], base)
    finally:
        _cabi.set_tuning("nvtx", 0)
    assert _cabi.get_tuning("nvtx") == 0


def test_sampler_cuda_graph_equals_eager(dev):
    """sample_dddm(cuda_graph=True): same Philox stream, same draws, same result as the eager loop."""
    import ddm_b200

    model = MixModel().to(dev)
    outs = []
    for graph in (False, True):
        torch.manual_seed(123)
        outs.append(ddm_b200.sample_dddm(model, n_samples=64, steps=7, eps_churn=0.7, device=str(dev), data_shape=(3, 8, 8),
                                         cuda_graph=graph))
        tail = torch.randn(4, device=dev)  # the generator must end at the same position
        outs.append(tail)
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3])


def test_sampler_graph_cache_follows_parameter_storage(dev):
    """A cached sampler graph must not outlive the parameter storage it captured (model.to(dtype) re-allocates)."""
    import ddm_b200
    from ddm_b200 import sampling

    model = MixModel().to(dev)
    torch.manual_seed(7)
    a = ddm_b200.sample_dddm(model, n_samples=16, steps=3, device=str(dev), data_shape=(4,), cuda_graph=True)
    n_graphs = len(sampling._graph_cache[model])
    with torch.no_grad():
        model.a.mul_(0.5)  # in-place update: same storage, same graph, new values
    torch.manual_seed(7)
    b = ddm_b200.sample_dddm(model, n_samples=16, steps=3, device=str(dev), data_shape=(4,), cuda_graph=True)
    assert len(sampling._graph_cache[model]) == n_graphs and not torch.equal(a, b)
    torch.manual_seed(7)
    ref = ddm_b200.sample_dddm(model, n_samples=16, steps=3, device=str(dev), data_shape=(4,))
    assert torch.equal(b, ref)
    model.double().float()  # re-allocates every parameter
    torch.manual_seed(7)
    c = ddm_b200.sample_dddm(model, n_samples=16, steps=3, device=str(dev), data_shape=(4,), cuda_graph=True)
    assert len(sampling._graph_cache[model]) == n_graphs + 1 and torch.equal(c, ref)


def test_toy_gmm_trains_end_to_end(dev):
    """BASELINE config 1 (2-D bimodal GMM, DDDMMLP, batch 512, m=8, beta=0.1, Adam 2e-3, 20 sampling steps) through the
    kernels.  The reference's own flow run on CPU for the same 1500 steps gives MMD^2(sigma=1) = 0.043 / 0.26 / 0.30 /
    0.23 for seeds 42 / 1 / 2 / 3 (from 1.33 untrained; run_example.py trains 5000-10000 steps): the kernels must land
    in the same range, on both modes."""
    from tools.toy_gmm_e2e import run

    r = run(steps=1500, seed=42, dev=str(dev), log_every=1500)
    assert r["mmd2_untrained"] > 1.0 and r["mmd2_rbf_sigma1"] < 0.45, r
    assert 0.3 < r["fraction_left_mode"] < 0.7 and r["fraction_within_3sigma_of_a_mode"] > 0.6, r
    h = r["history"][-1]
    assert 0.35 < h["loss"] < 0.55 and 0.85 < h["confidence"] < 1.05 and 0.75 < h["interaction"] < 0.95, h


def test_chunked_sampler_equals_per_chunk_calls(dev):
    """sample_dddm_chunked (evaluation-sized sample counts) == the concatenation of sample_dddm calls on the same RNG
    stream, ragged last chunk included; the full chunks share one cached graph."""
    import ddm_b200
    from ddm_b200 import sampling

    model = MixModel().to(dev)
    torch.manual_seed(5)
    got = ddm_b200.sample_dddm_chunked(model, 70, steps=4, device=str(dev), data_shape=(3, 4, 4), chunk=32)
    torch.manual_seed(5)
    want = torch.cat([ddm_b200.sample_dddm(model, n, steps=4, device=str(dev), data_shape=(3, 4, 4)) for n in (32, 32, 6)])
    assert got.shape == (70, 3, 4, 4) and torch.equal(got, want)
    assert len(sampling._graph_cache[model]) == 2  # one graph for the 32-sample chunks, one for the ragged tail
    assert ddm_b200.sample_dddm_chunked(model, 0, steps=2, device=str(dev), data_shape=(2,)).shape == (0, 2)


def test_trainer_cuda_graph_equals_eager(dev, monkeypatch):
    """launcher.Trainer: the whole-step CUDA graph (one GPU) and the two-graph form used with several ranks give the
    same parameters as the eager step after a few optimisation steps from the same seed — the warm-up steps of the
    capture are rolled back (weights, optimizer state, RNG position).  The very first bf16 run of a process picks
    its cuBLAS kernels cold and differs from every later run by rounding, so a throw-away eager run comes first."""
    from ddm_b200 import launcher

    flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    base = ["--synthetic", "--depth", "2", "--embed-dim", "96", "--heads", "3", "--batch", "8", "--m", "4", "--lr", "1e-3"]
    try:
        for precision, tol in (("bf16", 1e-3), ("fp32", 1e-4)):
            results = {}
            for mode in ("warmup", "eager", "graph", "split"):
                monkeypatch.setenv("DDDM_SPLIT_GRAPH", "1" if mode == "split" else "0")
                args = launcher.build_parser().parse_args(base + ["--precision", precision] +
                                                          (["--cuda-graph"] if mode in ("graph", "split") else ["--no-cuda-graph"]))
                tr = launcher.Trainer(args, dev, 1)
                assert tr.use_graph == (mode in ("graph", "split")) and tr.split_graph == (mode == "split")
                start = tr.flat_master.clone()
                losses = [tr.step(tr.synthetic_batch())["loss"] for _ in range(4)]
                results[mode] = (tr.flat_master - start, losses, torch.rand(3, device=dev))
            ref_d, ref_l, ref_r = results["eager"]
            assert float(ref_d.abs().max()) > 1e-3  # the optimizer moved the weights
            for mode in ("graph", "split"):
                d, l, r = results[mode]
                assert torch.equal(r, ref_r), (precision, mode, "RNG position differs")
                assert max(abs(a - b) for a, b in zip(l, ref_l)) <= 1e-5 * max(abs(b) for b in ref_l), (precision, mode, l, ref_l)
                rel = float((d - ref_d).norm() / ref_d.norm())
                assert rel <= tol, (precision, mode, rel)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = flags


def test_second_device_in_one_process(dev):
    """Host-side caches (shared-memory opt-in, SM count) are per device: a process that used GPU 0 can run the
    110 KB-tile kernels on GPU 1 — through the custom ops and through a host session."""
    from ddm_b200 import _cabi, ops

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    L = _cabi.lib()
    gen = torch.Generator().manual_seed(11)
    x0 = torch.randn(16, 3072, generator=gen).clamp(-1, 1)
    xh = x0[:, None] + 0.05 * torch.randn(16, 8, 3072, generator=gen)
    res = []
    for d in (0, 1, 0):
        device = torch.device("cuda", d)
        w = torch.tensor([8.0], device=device)
        out, grad = ops.energy_fused(xh.to(device), x0.to(device), w, 1.0 / 16, 0.1, 1.0, True)
        o32, g32 = ops.energy_fused(xh.to(device).repeat(1, 4, 1), x0.to(device), w, 1.0 / 16, 0.1, 1.0, True)  # m = 32: blocked kernel
        torch.cuda.synchronize(device)
        res.append((out.cpu(), grad.cpu(), o32.cpu(), g32.cpu()))
    for r in res[1:]:
        assert all(torch.equal(a, b) for a, b in zip(res[0], r))
    t = torch.rand(16, generator=gen)
    hx, h0, ht = xh.contiguous().pin_memory(), x0.pin_memory(), t.pin_memory()
    outs = []
    for d in (0, 1):
        s = L.dddm_session_create(16, 8, 3072, 0, d)
        assert s
        o = torch.empty(4).pin_memory()
        _cabi.check(L.dddm_session_step_host(s, hx.data_ptr(), h0.data_ptr(), ht.data_ptr(), 0.0, 0.1, 1.0, None,
                                             o.data_ptr()))
        L.dddm_session_destroy(s)
        outs.append(o.clone())
    assert torch.equal(outs[0], outs[1])
    torch.cuda.set_device(dev)
