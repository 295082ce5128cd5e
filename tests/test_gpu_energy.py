"""GPU parity of the energy-score kernels (K1 fused, K1b split), called through the C ABI via the
custom ops, against the CPU oracle and the golden fixtures recorded from the reference.

Tolerances (BASELINE.json north_star): fp32 loss and gradients within 1e-5 relative, bf16 within
1e-2 — relative to the largest magnitude of the reference tensor (normwise), measured against the
fp64 oracle evaluated on the SAME (fp32- or bf16-rounded) inputs.
"""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

FP32_REL = 1e-5
BF16_REL = 1e-2


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def _names(g):
    return [str(n) for n in g["names"]]


@pytest.fixture(scope="module")
def dev(cuda_device):
    import ddm_b200  # noqa: F401  (fails loudly if the library is missing)
    from ddm_b200 import _cabi

    _cabi.lib()
    yield cuda_device
    for k in ("energy.cluster", "energy.nv", "energy.variant", "energy.threads", "energy.ksmem"):
        _cabi.set_tuning(k, 0)
    _cabi.set_tuning("energy.pdl", 1)


def _fused(xhat, x0, w, beta, lam, want_grad=True):
    from ddm_b200 import ops

    wt = torch.tensor([w], dtype=torch.float32, device=xhat.device)
    out, grad = ops.energy_fused(xhat, x0, wt, 1.0, beta, lam, want_grad)
    return out.cpu().numpy().astype(np.float64), (grad.float().cpu().numpy() if want_grad else None)


def _check_case(xh_np, x0_np, beta, dev, dtype=torch.float32, rel=FP32_REL, lam=1.3, w=0.7):
    from ddm_b200 import ops

    xh = torch.from_numpy(np.asarray(xh_np, dtype=np.float32)).to(dev).to(dtype)
    x0 = torch.from_numpy(np.asarray(x0_np, dtype=np.float32)).to(dev).to(dtype)
    # oracle on exactly the values the kernel sees
    xh64, x064 = xh.double().cpu().numpy(), x0.double().cpu().numpy()
    loss, conf, inter, grad = oracle.energy_loss(xh64, x064, beta, lam, w)
    out, g = _fused(xh, x0, w, beta, lam)
    scale = max(abs(conf), abs(inter), 1e-30)
    assert abs(out[1] - conf) <= FP32_REL * scale, ("conf", out[1], conf)
    assert abs(out[2] - inter) <= FP32_REL * scale, ("inter", out[2], inter)
    assert abs(out[0] - loss) <= FP32_REL * max(abs(w) * scale, 1e-30) * 2, ("loss", out[0], loss)
    assert abs(out[3] - w) <= 1e-7
    gmax = np.max(np.abs(grad))
    if gmax > 0:
        assert _rel(g, grad) <= rel, ("grad", _rel(g, grad))
    else:
        assert not np.any(g)
    # split pair: forward values, saved distances, backward with independent upstream gradients
    o2, dist = ops.energy_terms_fwd(xh, x0, beta)
    o2 = o2.cpu().numpy().astype(np.float64)
    assert abs(o2[0] - conf) <= FP32_REL * scale and abs(o2[1] - inter) <= FP32_REL * scale
    gc = torch.tensor([0.37], device=dev)
    gi = torch.tensor([-1.9], device=dev)
    gx, gx0 = ops.energy_terms_bwd(xh, x0, dist, gc, gi, beta, True)
    ref_gx, ref_gx0 = oracle.energy_terms_grad(xh64, x064, beta, 0.37, -1.9, want_x0=True)
    # one common scale for both gradients (a vanishing grad_x0 next to an O(1) grad_xhat is "zero")
    gscale = max(np.max(np.abs(ref_gx)), np.max(np.abs(ref_gx0)), 1e-6)
    assert np.max(np.abs(gx.float().cpu().numpy() - ref_gx)) <= rel * gscale
    assert np.max(np.abs(gx0.float().cpu().numpy() - ref_gx0)) <= rel * gscale


def test_golden_cases_fp32(dev, golden_energy):
    g = golden_energy
    for n in _names(g):
        _check_case(g[f"{n}/xhat"], g[f"{n}/x0"], float(g[f"{n}/beta"]), dev)


def test_golden_values_against_reference_outputs(dev, golden_energy):
    """Directly against the numbers the reference produced (not via the oracle)."""
    from ddm_b200 import ops

    g = golden_energy
    for n in _names(g):
        xh = torch.from_numpy(g[f"{n}/xhat"]).to(dev)
        x0 = torch.from_numpy(g[f"{n}/x0"]).to(dev)
        beta = float(g[f"{n}/beta"])
        out, dist = ops.energy_terms_fwd(xh, x0, beta)
        out = out.cpu().numpy()
        ref = np.array([g[f"{n}/conf"], g[f"{n}/inter"]])
        assert np.max(np.abs(out - ref)) <= FP32_REL * np.max(np.abs(ref)), n
        one, zero = torch.ones(1, device=dev), torch.zeros(1, device=dev)
        gx, gx0 = ops.energy_terms_bwd(xh, x0, dist, one, zero, beta, True)
        gscale = np.max(np.abs(g[f"{n}/g_conf"]))  # grad_x0 = -sum_i of these rows: same scale (it may cancel to ~0)
        if gscale > 0:
            assert np.max(np.abs(gx.cpu().numpy() - g[f"{n}/g_conf"])) <= FP32_REL * gscale, n
            assert np.max(np.abs(gx0.cpu().numpy() - g[f"{n}/g_conf_x0"])) <= FP32_REL * gscale, n
        gx, _ = ops.energy_terms_bwd(xh, x0, dist, zero, one, beta, False)
        if np.max(np.abs(g[f"{n}/g_inter"])) > 0:
            assert _rel(gx.cpu().numpy(), g[f"{n}/g_inter"]) <= FP32_REL, n


def test_known_answers(dev, golden_energy):
    g = golden_energy
    out, grad = _fused(torch.from_numpy(g["kat1a/xhat"]).to(dev), torch.from_numpy(g["kat1a/x0"]).to(dev), 1.0, 2.0, 1.0)
    assert out[:3].tolist() == [-1.0, 1.0, 4.0] and grad.ravel().tolist() == [-1.0, 1.0]
    out, grad = _fused(torch.from_numpy(g["kat3/xhat"]).to(dev), torch.from_numpy(g["kat3/x0"]).to(dev), 1.0, 0.1, 1.0)
    assert abs(out[1] - 1e-12 ** 0.05) < 1e-6 and abs(out[2] - out[1]) < 1e-7 and not grad.any()
    assert np.isfinite(out).all()


def test_golden_cases_bf16(dev, golden_energy):
    g = golden_energy
    for n in _names(g):
        if n.startswith(("bf16grid", "early_B3_m8_D33", "late_B2_m8_D259", "early_B2_m16_D7", "dups")):
            _check_case(g[f"{n}/xhat"], g[f"{n}/x0"], float(g[f"{n}/beta"]), dev, dtype=torch.bfloat16, rel=BF16_REL)


def _synthetic(B, m, D, regime, seed=0):
    """SURVEY.md §8(d) generators: CIFAR-range x0, 'early' (independent) or 'late' (x0 + 0.05 noise) draws."""
    gen = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, D, generator=gen).clamp(-1, 1)
    if regime == "early":
        xh = torch.randn(B, m, D, generator=gen)
    else:
        xh = x0[:, None, :] + 0.05 * torch.randn(B, m, D, generator=gen)
    return xh, x0


@pytest.mark.parametrize("regime", ["early", "late"])
@pytest.mark.parametrize("beta", [0.1, 1.0, 2.0])
def test_headline_shape_fp32(dev, regime, beta):
    """BASELINE config 2 at full size: B=128, m=8, D=3072."""
    xh, x0 = _synthetic(128, 8, 3072, regime)
    _check_case(xh.numpy(), x0.numpy(), beta, dev)


@pytest.mark.parametrize("regime", ["early", "late"])
def test_headline_shape_bf16(dev, regime):
    xh, x0 = _synthetic(128, 8, 3072, regime, seed=1)
    _check_case(xh.numpy(), x0.numpy(), 0.1, dev, dtype=torch.bfloat16, rel=BF16_REL)


@pytest.mark.parametrize("m", [4, 8, 16, 32])
@pytest.mark.parametrize("D", [2, 3072, 12288])
def test_sweep_shapes(dev, m, D):
    """BASELINE config 3 grid (m x D), beta cycling over {0.1, 1.0, 2.0}; B reduced where the oracle is slow."""
    B = 128 if m * D <= 8 * 3072 else 16
    for k, beta in enumerate((0.1, 1.0, 2.0)):
        regime = "late" if (k + m) % 2 else "early"
        xh, x0 = _synthetic(B, m, D, regime, seed=m * 7 + k)
        _check_case(xh.numpy(), x0.numpy(), beta, dev)
        if k == 0 and D >= 8:
            _check_case(xh.numpy(), x0.numpy(), beta, dev, dtype=torch.bfloat16, rel=BF16_REL)


@pytest.mark.parametrize("m,D", [(2, 1), (3, 5), (5, 130), (7, 64), (8, 259), (8, 4100), (9, 33), (12, 100), (17, 68),
                                 (33, 40), (64, 36), (8, 96), (6, 8192)])
def test_ragged_shapes(dev, m, D):
    """Odd m, D not a multiple of the vector width (unaligned rows), slabs with empty trailing CTAs."""
    for B, beta in ((1, 0.1), (3, 1.5), (5, 2.0)):
        xh, x0 = _synthetic(B, m, D, "early", seed=B + m + D)
        _check_case(xh.numpy(), x0.numpy(), beta, dev)
    xh, x0 = _synthetic(2, m, D, "late", seed=99)
    _check_case(xh.numpy(), x0.numpy(), 0.1, dev, dtype=torch.bfloat16, rel=BF16_REL)


@pytest.mark.parametrize("B,m,D,regime,beta", [(128, 8, 3072, "late", 0.1), (128, 8, 3072, "early", 1.0), (5, 3, 40, "late", 2.0),
                                               (9, 8, 12288, "late", 0.1), (3, 5, 4104, "early", 0.1), (300, 8, 1024, "late", 1.0)])
def test_mixed_bf16_draws_fp32_data(dev, B, m, D, regime, beta):
    """Mixed entry point (bf16 xhat as a bf16 backbone writes it, fp32 x0, bf16 gradient): against the fp64 oracle on
    the bf16-rounded draws and the EXACT fp32 data — and strictly closer to it than the all-bf16 entry, which rounds x0."""
    from ddm_b200 import _cabi

    assert _cabi.lib().dddm_energy_fused_bf16_x0f32_supported(m, D) == 1
    xh, x0 = _synthetic(B, m, D, regime, seed=B + m)
    xh, x0 = xh.to(dev).to(torch.bfloat16), x0.to(dev)
    loss, conf, inter, grad = oracle.energy_loss(xh.double().cpu().numpy(), x0.double().cpu().numpy(), beta, 1.3, 0.7)
    out, g = _fused(xh, x0, 0.7, beta, 1.3)
    assert g.shape == tuple(xh.shape)
    scale = max(abs(conf), abs(inter))
    assert abs(out[1] - conf) <= FP32_REL * scale and abs(out[2] - inter) <= FP32_REL * scale  # values are fp32-exact
    assert abs(out[0] - loss) <= 2 * FP32_REL * 0.7 * scale
    assert _rel(g, grad) <= BF16_REL  # the gradient is rounded to bf16 on the way out
    if regime == "late" and D >= 1024:  # rounding x0 to bf16 (|x0| <= 1, draws 0.05 away) visibly moves the terms
        out_b, _ = _fused(xh, x0.to(torch.bfloat16), 0.7, beta, 1.3)
        assert abs(out_b[1] - conf) > 10 * abs(out[1] - conf)
    assert _cabi.lib().dddm_energy_fused_bf16_x0f32_supported(16, 3072) == 1  # m = 16 / 32: the tensor-core kernel
    assert _cabi.lib().dddm_energy_fused_bf16_x0f32_supported(24, 3072) == 0
    with pytest.raises(TypeError):
        _fused(xh.float(), x0.to(torch.bfloat16), 0.7, beta, 1.3)


def test_wave_kernel_plans(dev):
    """The single-wave register-resident kernel (variant 5, the path one launch of B <= #SMs rows takes): every
    (threads, vectors per thread, coefficient placement) plan, fp32 and bf16, m = 2..8, rows that do not fill the
    last vector slot, more rows than SMs (forced), fused and split-forward modes — against the fp64 oracle."""
    from ddm_b200 import _cabi, ops

    assert "wave<" not in _cabi.describe_energy(128, 8, 3072)  # opt-in plan
    seen = set()
    try:
        _cabi.set_tuning("energy.variant", 5)
        for dtype, rel, name in ((torch.float32, FP32_REL, "f32"), (torch.bfloat16, BF16_REL, "bf16")):
            vecw = 4 if name == "f32" else 8
            for threads, nv in ((128, 1), (128, 2), (128, 3), (256, 1), (256, 2), (256, 3), (384, 1), (384, 2)):
                for ksmem in (0, 1):  # (selects the m used below; the pass-2 form is always pair-major)
                    _cabi.set_tuning("energy.threads", threads)
                    _cabi.set_tuning("energy.nv", nv)
                    _cabi.set_tuning("energy.ksmem", ksmem)
                    # full tile, and a row that ends inside the last vector slot of some threads
                    for D in (threads * nv * vecw, threads * nv * vecw - 5 * vecw):
                        m = 8 if ksmem == 0 else 2 + (threads // 128 + nv) % 7
                        B = 5
                        xh, x0 = _synthetic(B, m, D, "late" if nv % 2 else "early", seed=threads + nv)
                        beta = (0.1, 1.0, 2.0)[(nv + ksmem) % 3]
                        xh, x0 = xh.to(dev).to(dtype), x0.to(dev).to(dtype)
                        loss, conf, inter, grad = oracle.energy_loss(xh.double().cpu().numpy(), x0.double().cpu().numpy(),
                                                                     beta, 1.3, 0.7)
                        out, g = _fused(xh, x0, 0.7, beta, 1.3)
                        seen.add(_cabi.describe_energy(B, m, D, name))
                        scale = max(abs(conf), abs(inter))
                        assert abs(out[1] - conf) <= FP32_REL * scale and abs(out[2] - inter) <= FP32_REL * scale
                        assert abs(out[0] - loss) <= 2 * FP32_REL * 0.7 * scale
                        assert _rel(g, grad) <= rel, (name, threads, nv, ksmem, D, _rel(g, grad))
                        o2, dist = ops.energy_terms_fwd(xh, x0, beta)  # split forward on the same kernel
                        o2 = o2.cpu().numpy().astype(np.float64)
                        assert abs(o2[0] - conf) <= FP32_REL * scale and abs(o2[1] - inter) <= FP32_REL * scale
        # more rows than SMs (several waves of one CTA per SM) and repeated launches on one workspace
        for k in ("energy.threads", "energy.nv", "energy.ksmem"):
            _cabi.set_tuning(k, 0)
        xh, x0 = _synthetic(333, 8, 1024, "late", seed=3)
        xh, x0 = xh.to(dev), x0.to(dev)
        loss, conf, inter, grad = oracle.energy_loss(xh.double().cpu().numpy(), x0.double().cpu().numpy(), 0.1, 1.0, 0.5)
        for _ in range(3):
            out, g = _fused(xh, x0, 0.5, 0.1, 1.0)
            assert "wave<" in _cabi.describe_energy(333, 8, 1024)
            assert abs(out[0] - loss) <= 2e-5 * abs(conf)
            assert _rel(g, grad) <= FP32_REL
        # forward only (no gradient buffer)
        out, _ = _fused(xh, x0, 0.5, 0.1, 1.0, want_grad=False)
        assert abs(out[0] - loss) <= 2e-5 * abs(conf)
    finally:
        for k in ("energy.variant", "energy.threads", "energy.nv", "energy.ksmem"):
            _cabi.set_tuning(k, 0)
    assert len(seen) >= 24, len(seen)


def test_pipe_kernel_plans(dev):
    """The row-pipelined cluster kernel (variant 6: C CTAs own C rows, one column slab of each; pass 2 of a row runs
    while the next rows are still in flight): every cluster size, thread count, step width and request window, fp32,
    bf16 and the mixed entry, row counts that do not fill the last cluster, slabs that do not divide D, fused and
    split-forward modes, repeated launches on one workspace — against the fp64 oracle."""
    from ddm_b200 import _cabi, ops

    seen = set()
    try:
        _cabi.set_tuning("energy.variant", 6)
        case = 0
        for dtype, rel, name in ((torch.float32, FP32_REL, "f32"), (torch.bfloat16, BF16_REL, "bf16")):
            for cluster in (2, 4, 8):
                for threads, cols in ((128, 4), (256, 2), (96, 4), (64, 2)):
                    for window in (0, 1, 3):
                        case += 1
                        _cabi.set_tuning("energy.cluster", cluster)
                        _cabi.set_tuning("energy.threads", threads)
                        _cabi.set_tuning("energy.cols", cols)
                        _cabi.set_tuning("energy.window", window)
                        m = 8 if case % 3 else 4
                        B = (5, 16, 9, 1)[case % 4]
                        D = (3072, 1000, 3072 - 40, 64)[case % 4]
                        beta = (0.1, 1.0, 2.0)[case % 3]
                        xh, x0 = _synthetic(B, m, D, "late" if case % 2 else "early", seed=case)
                        xh, x0 = xh.to(dev).to(dtype), x0.to(dev).to(dtype)
                        loss, conf, inter, grad = oracle.energy_loss(xh.double().cpu().numpy(), x0.double().cpu().numpy(),
                                                                     beta, 1.3, 0.7)
                        desc = _cabi.describe_energy(B, m, D, name)
                        assert desc.startswith("pipe<"), desc
                        seen.add(desc)
                        for _ in range(2):
                            out, g = _fused(xh, x0, 0.7, beta, 1.3)
                            scale = max(abs(conf), abs(inter))
                            assert abs(out[1] - conf) <= FP32_REL * scale and abs(out[2] - inter) <= FP32_REL * scale, desc
                            assert abs(out[0] - loss) <= 2 * FP32_REL * 0.7 * scale, desc
                            assert _rel(g, grad) <= rel, (desc, _rel(g, grad))
                        o2, dist = ops.energy_terms_fwd(xh, x0, beta)  # split forward on the same kernel
                        o2 = o2.cpu().numpy().astype(np.float64)
                        assert abs(o2[0] - conf) <= FP32_REL * scale and abs(o2[1] - inter) <= FP32_REL * scale
                        gc, gi = torch.tensor([0.37], device=dev), torch.tensor([-1.9], device=dev)
                        gx, _ = ops.energy_terms_bwd(xh, x0, dist, gc, gi, beta, False)  # consumes the saved distances
                        ref_gx, _ = oracle.energy_terms_grad(xh.double().cpu().numpy(), x0.double().cpu().numpy(), beta,
                                                             0.37, -1.9, want_x0=True)
                        assert np.max(np.abs(gx.float().cpu().numpy() - ref_gx)) <= rel * max(np.max(np.abs(ref_gx)), 1e-6)
        # BASELINE config 2 at full size, default plan, fp32 + bf16 + mixed entry (bf16 draws, fp32 data)
        for k in ("energy.cluster", "energy.threads", "energy.cols", "energy.window"):
            _cabi.set_tuning(k, 0)
        xh, x0 = _synthetic(128, 8, 3072, "late", seed=11)
        for dtype, rel in ((torch.float32, FP32_REL), (torch.bfloat16, BF16_REL)):
            a, c = xh.to(dev).to(dtype), x0.to(dev).to(dtype)
            loss, conf, inter, grad = oracle.energy_loss(a.double().cpu().numpy(), c.double().cpu().numpy(), 0.1, 1.0, 0.5)
            out, g = _fused(a, c, 0.5, 0.1, 1.0)
            assert abs(out[0] - loss) <= 2e-5 * abs(conf)
            assert _rel(g, grad) <= rel
            out, _ = _fused(a, c, 0.5, 0.1, 1.0, want_grad=False)
            assert abs(out[0] - loss) <= 2e-5 * abs(conf)
        a, c = xh.to(dev).to(torch.bfloat16), x0.to(dev)
        loss, conf, inter, grad = oracle.energy_loss(a.double().cpu().numpy(), c.double().cpu().numpy(), 0.1, 1.0, 0.5)
        out, g = _fused(a, c, 0.5, 0.1, 1.0)
        assert abs(out[0] - loss) <= 2e-5 * abs(conf)
        assert _rel(g, grad) <= BF16_REL
    finally:
        for k in ("energy.variant", "energy.cluster", "energy.threads", "energy.cols", "energy.window"):
            _cabi.set_tuning(k, 0)
    assert len(seen) >= 40, len(seen)


def _tc_inputs(B, m, D, regime, seed):
    """late / early as _synthetic; 'mixed' = spread draws with a few near- and exact duplicates and one draw on top of
    x0; 'dups' = identical draws (a zero-initialised output layer), one equal to x0, one a hair off."""
    gen = torch.Generator().manual_seed(seed)
    if regime in ("late", "early"):
        return _synthetic(B, m, D, regime, seed)
    x0 = torch.randn(B, D, generator=gen).clamp(-1, 1)
    if regime == "dups":
        xh = x0[:, None, :].repeat(1, m, 1) * 0.5
        xh[:, 1] = x0
        xh[:, 3] += 1e-3 * torch.randn(B, D, generator=gen)
    else:
        xh = x0[:, None, :] + 0.3 * torch.randn(B, m, D, generator=gen)
        xh[:, 5] = xh[:, 2] + 2e-3 * torch.randn(B, D, generator=gen)
        xh[:, 7] = xh[:, 2]
        xh[:, 9] = x0 + 1e-3 * torch.randn(B, D, generator=gen)
    return xh, x0


@pytest.mark.parametrize("m", [16, 32])
def test_tensor_core_kernel(dev, m):
    """The tcgen05 kernel (variant 7; the default for m = 32 bf16 and for the mixed entry): Gram of the centred draws +
    coefficient mixing on the tensor cores, against the fp64 oracle on the same bf16 inputs — every regime (incl. exact
    and near duplicates, which take the direct-difference fallbacks), all betas, bf16 and fp32 x0, more rows than SMs,
    narrow and wide rows, forward-only, saved distances of the split forward, determinism."""
    from ddm_b200 import _cabi, ops

    assert _cabi.describe_energy(128, 32, 3072, "bf16").startswith("tc<")  # the default plan at BASELINE config 3
    assert _cabi.describe_energy(128, 16, 3072, "bf16").startswith("tc<") and _cabi.describe_energy(128, 16, 3072).startswith("blk<")
    try:
        _cabi.set_tuning("energy.variant", 7)
        case = 0
        for B, D in ((3, 128), (8, 3072), (150, 768), (2, 12288)):
            for regime in ("late", "early", "mixed", "dups"):
                case += 1
                beta = (0.1, 1.0, 2.0)[case % 3]
                xh, x0 = _tc_inputs(B, m, D, regime, seed=case)
                xh = xh.to(dev).to(torch.bfloat16)
                for x0_dtype in (torch.bfloat16, torch.float32):
                    c = x0.to(dev).to(x0_dtype)
                    assert _cabi.describe_energy(B, m, D, "bf16").startswith("tc<")
                    xh64, x064 = xh.double().cpu().numpy(), c.double().cpu().numpy()
                    loss, conf, inter, grad = oracle.energy_loss(xh64, x064, beta, 1.3, 0.7)
                    out, g = _fused(xh, c, 0.7, beta, 1.3)
                    scale = max(abs(conf), abs(inter), 1e-30)
                    tag = (m, B, D, regime, beta, str(x0_dtype))
                    assert abs(out[1] - conf) <= 1e-3 * scale and abs(out[2] - inter) <= 1e-3 * scale, tag
                    assert abs(out[0] - loss) <= 1e-3 * 0.7 * scale, tag
                    assert _rel(g, grad) <= BF16_REL, (tag, _rel(g, grad))
                    out2, g2 = _fused(xh, c, 0.7, beta, 1.3)  # same bits on a second launch
                    assert np.array_equal(out, out2) and np.array_equal(g, g2), tag
                    o3, _ = _fused(xh, c, 0.7, beta, 1.3, want_grad=False)
                    assert np.array_equal(o3, out), tag
                # split forward: saved squared distances (m confinement, then pairs i < j row-major)
                c = x0.to(dev).to(torch.bfloat16)
                o2, dist = ops.energy_terms_fwd(xh, c, beta)
                xh64, x064 = xh.double().cpu().numpy(), c.double().cpu().numpy()
                iu = np.triu_indices(m, 1)
                d_ref = np.concatenate([((xh64 - x064[:, None]) ** 2).sum(-1),
                                        ((xh64[:, :, None] - xh64[:, None]) ** 2).sum(-1)[:, iu[0], iu[1]]], axis=1)
                d = dist.cpu().numpy().astype(np.float64)
                ok = np.abs(d - d_ref) <= 2e-3 * d_ref + 1e-12
                assert ok.all(), (m, B, D, regime, float(np.max(np.abs(d - d_ref) / np.maximum(d_ref, 1e-30))))
                # split backward on the same kernel (distances from `dist`, two upstream gradients, grad_x0 = -sum_i grad_i)
                gc, gi = torch.tensor([0.37], device=dev), torch.tensor([-1.9], device=dev)
                gx, gx0 = ops.energy_terms_bwd(xh, c, dist, gc, gi, beta, True)
                ref_gx, ref_gx0 = oracle.energy_terms_grad(xh64, x064, beta, 0.37, -1.9, want_x0=True)
                gscale = max(np.max(np.abs(ref_gx)), np.max(np.abs(ref_gx0)), 1e-6)
                assert np.max(np.abs(gx.float().cpu().numpy() - ref_gx)) <= BF16_REL * gscale, (m, B, D, regime, "bwd grad_xhat")
                assert np.max(np.abs(gx0.float().cpu().numpy() - ref_gx0)) <= BF16_REL * gscale, (m, B, D, regime, "bwd grad_x0")
                gx_only, none = ops.energy_terms_bwd(xh, c, dist, gc, gi, beta, False)
                assert torch.equal(gx_only, gx)
    finally:
        _cabi.set_tuning("energy.variant", 0)


@pytest.mark.parametrize("beta", [0.1, 1.0, 2.0])
def test_centred_pass2_and_direct_fallback(dev, beta):
    """Pass 2 of the TMA-staged kernel takes the centred form (g_i = (c_i + sum k_ij) z_i - sum k_ij z_j, z = x - x0) on
    rows whose draws are spread and the direct-difference form on rows with near- or exact duplicates; both must hold
    the fp32 tolerance, inside one launch that mixes such rows."""
    m, D = 8, 3072
    parts = [_tc_inputs(6, 16, D, r, seed=31 + k) for k, r in enumerate(("late", "mixed", "early", "dups"))]
    xh = torch.cat([p[0][:, :m] for p in parts])  # 'mixed' keeps its near-duplicates among the first 8 draws: (2, 5), (2, 7)
    x0 = torch.cat([p[1] for p in parts])
    _check_case(xh.numpy(), x0.numpy(), beta, dev)
    _check_case(xh.numpy(), x0.numpy(), beta, dev, dtype=torch.bfloat16, rel=BF16_REL)


def test_kernel_variants_agree(dev):
    """Every launch plan (cluster size, vectors per thread, register vs smem-tile variant) is the same function."""
    from ddm_b200 import _cabi

    xh, x0 = _synthetic(16, 8, 3072, "late", seed=5)
    xh, x0 = xh.to(dev), x0.to(dev)
    loss, conf, inter, grad = oracle.energy_loss(xh.double().cpu().numpy(), x0.double().cpu().numpy(), 0.1, 1.0, 0.5)
    seen = set()
    try:
        for variant in (1, 2, 3):
            for cluster in (1, 2, 4, 8):
                for knob in (1, 2):
                    _cabi.set_tuning("energy.variant", variant)
                    _cabi.set_tuning("energy.cluster", cluster)
                    _cabi.set_tuning("energy.nv", knob)
                    _cabi.set_tuning("energy.threads", 32 * knob)
                    nv = knob
                    for loader in ((1, 2) if variant == 3 else (0,)):  # TMA bulk copies / cp.async commit groups
                        _cabi.set_tuning("energy.loader", loader)
                        try:
                            out, g = _fused(xh, x0, 0.5, 0.1, 1.0)
                        except _cabi.DDDMError as e:
                            assert e.status == -4  # this plan does not cover the shape (e.g. register tile too wide)
                            continue
                        seen.add(_cabi.describe_energy(16, 8, 3072))
                        assert abs(out[0] - loss) <= 2e-5 * abs(conf), (variant, cluster, nv, loader)
                        assert _rel(g, grad) <= FP32_REL, (variant, cluster, nv, loader, _rel(g, grad))
        # the 8-compute-warp build of the TMA-staged kernel (one CTA per SM; opt-in: "energy.threads" = 256)
        _cabi.set_tuning("energy.variant", 3)
        _cabi.set_tuning("energy.cluster", 0)
        _cabi.set_tuning("energy.loader", 0)
        for nv in (1, 2):
            _cabi.set_tuning("energy.threads", 256)
            _cabi.set_tuning("energy.nv", nv)
            assert "threads=256" in _cabi.describe_energy(16, 8, 3072)
            out, g = _fused(xh, x0, 0.5, 0.1, 1.0)
            seen.add(_cabi.describe_energy(16, 8, 3072))
            assert abs(out[0] - loss) <= 2e-5 * abs(conf) and _rel(g, grad) <= FP32_REL, ("threads=256", nv, _rel(g, grad))
        _cabi.set_tuning("energy.threads", 0)
        for pdl in (0, 1):
            _cabi.set_tuning("energy.variant", 0)
            _cabi.set_tuning("energy.cluster", 0)
            _cabi.set_tuning("energy.nv", 0)
            _cabi.set_tuning("energy.pdl", pdl)
            out, g = _fused(xh, x0, 0.5, 0.1, 1.0)
            assert _rel(g, grad) <= FP32_REL
    finally:
        for k in ("energy.cluster", "energy.nv", "energy.variant", "energy.threads", "energy.loader"):
            _cabi.set_tuning(k, 0)
        _cabi.set_tuning("energy.pdl", 1)
    assert len(seen) >= 16, seen


def test_finish_protocols_bit_identical(dev):
    """TMA-staged kernel: the cross-row sum by polled row slots (default) and by the arrival ticket give bit-identical
    sums and gradients (same arithmetic, same summation order) — on one workspace used alternately by both protocols,
    for more rows than one wave holds, ragged D, D-split clusters, fp32 / bf16 / mixed inputs."""
    from ddm_b200 import _cabi

    try:
        for B, m, D, cluster, dtype, x0f32 in ((128, 8, 3072, 0, torch.float32, False), (700, 8, 3072, 0, torch.float32, False),
                                               (5, 7, 1000, 0, torch.float32, False), (33, 8, 3072 - 40, 2, torch.float32, False),
                                               (128, 8, 3072, 0, torch.bfloat16, False), (128, 8, 3072, 0, torch.bfloat16, True),
                                               (9, 5, 264, 4, torch.bfloat16, False), (1, 2, 8, 0, torch.float32, False)):
            xh, x0 = _synthetic(B, m, D, "late", seed=B + D)
            xh, x0 = xh.to(dev).to(dtype), x0.to(dev).to(torch.float32 if x0f32 else dtype)
            _cabi.set_tuning("energy.variant", 3)
            _cabi.set_tuning("energy.cluster", cluster)
            assert _cabi.describe_energy(B, m, D, "f32" if dtype == torch.float32 else "bf16").startswith("smem<")
            ref_out = ref_g = None
            for finish in (1, 2, 0, 1, 1, 2, 2):
                _cabi.set_tuning("energy.finish", finish)
                out, g = _fused(xh, x0, 0.7, 0.1, 1.3)
                assert np.all(np.isfinite(out))
                if ref_out is None:
                    ref_out, ref_g = out, g
                    loss, conf, inter, grad = oracle.energy_loss(xh.double().cpu().numpy(), x0.double().cpu().numpy(), 0.1, 1.3, 0.7)
                    assert abs(out[1] - conf) <= FP32_REL * max(abs(conf), abs(inter))
                    assert _rel(g, grad) <= (FP32_REL if dtype == torch.float32 else BF16_REL)
                else:
                    assert np.array_equal(out, ref_out), (B, m, D, finish, out, ref_out)
                    assert np.array_equal(g, ref_g), (B, m, D, finish)
    finally:
        for k in ("energy.variant", "energy.cluster", "energy.finish"):
            _cabi.set_tuning(k, 0)


def test_workspace_reset_recovers_a_corrupted_workspace(dev):
    """A workspace left half-written by an aborted launch may give wrong sums (never a hang); dddm_energy_workspace_reset makes it usable again
    (both cross-row protocols share the slots: TMA-staged kernel = polled slots, register kernel = arrival ticket)."""
    from ddm_b200 import _cabi

    L = _cabi.lib()
    stream = torch.cuda.current_stream().cuda_stream
    for B, m, D in ((64, 8, 3072), (9, 4, 6)):  # TMA-staged kernel / register-resident kernel (unaligned rows)
        xh, x0 = _synthetic(B, m, D, "late", seed=5)
        xh, x0 = xh.to(dev), x0.to(dev)
        w = torch.full((1,), 0.5 * B, device=dev)
        out, grad = torch.zeros(4, device=dev), torch.empty_like(xh)
        ws = torch.zeros(L.dddm_energy_workspace_bytes(B, m), dtype=torch.uint8, device=dev)

        def run():
            out.zero_()
            _cabi.check(L.dddm_energy_fused_f32(xh.data_ptr(), x0.data_ptr(), w.data_ptr(), 1.0 / B, grad.data_ptr(), out.data_ptr(),
                                                ws.data_ptr(), B, m, D, 0.1, 1.0, stream))
            torch.cuda.synchronize()
            return out.clone()

        good = run()
        assert torch.equal(run(), good)
        ws[16:].view(torch.int64)[: B // 2] = -(2**62)  # half of the rows look "arrived" (bit 63) with garbage sums
        ws[:4].view(torch.int32)[0] = 3                   # and the ticket is left in the middle of a count
        run()  # sums may be anything (which protocol step sees the garbage first is a race) — but the launch must end
        _cabi.check(L.dddm_energy_workspace_reset(ws.data_ptr(), B, m, stream))
        assert torch.equal(run(), good)
        assert torch.equal(run(), good)
    assert L.dddm_energy_workspace_reset(None, 4, 8, stream) == -1


@pytest.mark.parametrize("m", [16, 24, 32])
def test_blocked_kernel_plans_agree(dev, m):
    """m = 16 / 32: the blocked packed-fp32 kernel (variant 4) under every cluster size and thread count, and the
    generic chunked tile kernel (variant 2), against the oracle — fp32 and bf16, early and late regime, all betas."""
    from ddm_b200 import _cabi

    seen = set()
    try:
        for D, B in ((3072, 6), (1024, 5), (12288, 3), (64, 7)):
            for regime, beta in (("late", 0.1), ("early", 1.0), ("late", 2.0)):
                xh, x0 = _synthetic(B, m, D, regime, seed=m + D)
                for variant, cluster, threads in ((4, 0, 0), (4, 1, 0), (4, 2, 64), (4, 4, 0), (4, 8, 32), (2, 0, 0)):
                    _cabi.set_tuning("energy.variant", variant)
                    _cabi.set_tuning("energy.cluster", cluster)
                    _cabi.set_tuning("energy.threads", threads)
                    try:
                        _check_case(xh.numpy(), x0.numpy(), beta, dev)
                        if beta == 0.1:
                            _check_case(xh.numpy(), x0.numpy(), beta, dev, dtype=torch.bfloat16, rel=BF16_REL)
                    except _cabi.DDDMError as e:
                        assert e.status == -4, e  # plan does not cover the shape (tile too large for one CTA)
                        continue
                    seen.add(_cabi.describe_energy(B, m, D))
    finally:
        for k in ("energy.cluster", "energy.variant", "energy.threads"):
            _cabi.set_tuning(k, 0)
    assert sum(d.startswith("blk<") for d in seen) >= 8 and any(d.startswith("tile<") for d in seen), seen


def test_deterministic_and_workspace_reuse(dev):
    xh, x0 = _synthetic(128, 8, 3072, "early", seed=3)
    xh, x0 = xh.to(dev), x0.to(dev)
    first = _fused(xh, x0, 0.4, 0.1, 1.0)
    for _ in range(5):
        again = _fused(xh, x0, 0.4, 0.1, 1.0)
        assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1])


def test_properties_full_size(dev):
    """Size-independent properties at the headline shape (no oracle needed)."""
    from ddm_b200 import ops

    B, m, D = 128, 8, 3072
    xh, x0 = _synthetic(B, m, D, "late", seed=11)
    xh, x0 = xh.to(dev), x0.to(dev)
    base, gbase = _fused(xh, x0, 1.0, 1.0, 1.0)
    # translation invariance: distances are unchanged when every point moves by the same vector
    shift = torch.randn(B, 1, D, device=dev) * 0.25
    out, g = _fused(xh + shift, x0 + shift[:, 0], 1.0, 1.0, 1.0)
    assert np.allclose(out[:3], base[:3], rtol=2e-5)
    assert _rel(g, gbase) <= 2e-4  # inputs themselves were re-rounded by the shift
    # permuting the draws permutes the gradient rows and leaves the scalars unchanged
    perm = torch.randperm(m)
    out, g = _fused(xh[:, perm].contiguous(), x0, 1.0, 1.0, 1.0)
    assert np.allclose(out[:3], base[:3], rtol=1e-6)
    assert _rel(g, gbase[:, perm.numpy()]) <= 1e-6
    # homogeneity for beta = 2: f(a x) = a^2 f(x), grad(a x) = a grad(x)
    b2, g2 = _fused(xh, x0, 1.0, 2.0, 1.0)
    out, g = _fused(2.0 * xh, 2.0 * x0, 1.0, 2.0, 1.0)
    assert np.allclose(out[:3], 4.0 * b2[:3], rtol=1e-6) and _rel(g, 2.0 * g2) <= 1e-6
    # linearity in the weight, and loss == W * (conf - lam/(2(m-1)) inter)
    out, g = _fused(xh, x0, 0.25, 1.0, 3.0)
    assert abs(out[0] - 0.25 * (out[1] - 3.0 / (2 * (m - 1)) * out[2])) <= 1e-6 * abs(out[1])
    # rows are independent: the batch mean is the mean of the two half-batch means
    o1, _ = _fused(xh[:64].contiguous(), x0[:64].contiguous(), 1.0, 1.0, 1.0)
    o2, _ = _fused(xh[64:].contiguous(), x0[64:].contiguous(), 1.0, 1.0, 1.0)
    assert np.allclose(0.5 * (o1[1:3] + o2[1:3]), base[1:3], rtol=1e-6)
    # sum over draws of the interaction gradient vanishes (pair terms are antisymmetric)
    o, dist = ops.energy_terms_fwd(xh, x0, 1.0)
    gi, _ = ops.energy_terms_bwd(xh, x0, dist, torch.zeros(1, device=dev), torch.ones(1, device=dev), 1.0, False)
    assert float(gi.sum(dim=1).abs().max()) <= 1e-6 * float(gi.abs().max()) * m


def test_autograd_fused_and_split(dev):
    from ddm_b200 import generalized_energy_terms, ops

    xh, x0 = _synthetic(8, 8, 256, "early", seed=21)
    xh, x0 = xh.to(dev), x0.to(dev)
    xh64, x064 = xh.double().cpu().numpy(), x0.double().cpu().numpy()
    w = torch.tensor([0.6], device=dev)
    # fused op: loss.backward() hands the kernel's gradient over; upstream scale != 1 rescales it
    for upstream in (1.0, 2.5):
        a = xh.clone().requires_grad_(True)
        out, _ = ops.energy_fused(a, x0, w, 1.0, 0.1, 1.0, True)
        (out[0] * upstream).backward()
        _, _, _, ref = oracle.energy_loss(xh64, x064, 0.1, 1.0, 0.6)
        assert _rel(a.grad.cpu().numpy(), upstream * ref) <= FP32_REL
    a = xh.clone().requires_grad_(True)
    out, _ = ops.energy_fused(a, x0, w, 1.0, 0.1, 1.0, True)
    out[0].backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="already consumed"):
        out[0].backward()
    # split path: arbitrary function of (conf, inter), gradient w.r.t. both inputs
    a = xh.clone().requires_grad_(True)
    c = x0.clone().requires_grad_(True)
    conf, inter = generalized_energy_terms(a, c, 1.5, 1.0)
    (3.0 * conf - 0.5 * inter).backward()
    ref_gx, ref_gx0 = oracle.energy_terms_grad(xh64, x064, 1.5, 3.0, -0.5, want_x0=True)
    assert _rel(a.grad.cpu().numpy(), ref_gx) <= FP32_REL and _rel(c.grad.cpu().numpy(), ref_gx0) <= FP32_REL
    assert conf.dim() == 0 and inter.dim() == 0 and conf.dtype == torch.float32
    with pytest.raises(ValueError):
        generalized_energy_terms(xh[:, :1], x0, 1.0, 1.0)


def test_errors_are_loud(dev):
    from ddm_b200 import generalized_energy_terms, ops

    with pytest.raises(RuntimeError, match="CUDA-only"):
        generalized_energy_terms(torch.zeros(2, 2, 2), torch.zeros(2, 2), 1.0, 1.0)
    with pytest.raises(TypeError):
        ops.energy_terms_fwd(torch.zeros(2, 2, 2, device=dev, dtype=torch.float64),
                             torch.zeros(2, 2, device=dev, dtype=torch.float64), 1.0)
    with pytest.raises(ValueError):
        ops.energy_terms_fwd(torch.zeros(2, 2, 3, device=dev), torch.zeros(2, 2, device=dev), 1.0)


@pytest.mark.parametrize("m,D,B", [(8, 3072, 1500), (4, 768, 5000), (16, 1024, 700), (8, 2, 20000)])
def test_many_rows_multi_wave(dev, m, D, B):
    """More rows than resident CTAs (several waves of the grid, last-arriver reduction over thousands of rows):
    the batch means must equal the mean of per-slice means, the gradient of a slice must equal the slice's own
    gradient scaled by its share of the batch, and a sample of rows is checked against the oracle."""
    xh, x0 = _synthetic(B, m, D, "late", seed=B + m)
    xh, x0 = xh.to(dev), x0.to(dev)
    out, g = _fused(xh, x0, 0.6, 0.1, 1.0)
    cuts = [0, B // 3, B // 3 + 77, B]
    conf = inter = 0.0
    for a, b in zip(cuts[:-1], cuts[1:]):
        o, gs = _fused(xh[a:b].contiguous(), x0[a:b].contiguous(), 0.6, 0.1, 1.0)
        conf += o[1] * (b - a) / B
        inter += o[2] * (b - a) / B
        assert _rel(g[a:b], gs * (b - a) / B) <= 2e-6
    assert abs(conf - out[1]) <= 2e-6 * abs(out[1]) and abs(inter - out[2]) <= 2e-6 * abs(out[2])
    rows = np.linspace(0, B - 1, 9).astype(int)
    sub_h, sub_0 = xh[rows].double().cpu().numpy(), x0[rows].double().cpu().numpy()
    _, _, _, gref = oracle.energy_loss(sub_h, sub_0, 0.1, 1.0, 0.6)
    assert _rel(g[rows], gref * len(rows) / B) <= FP32_REL
