"""Pin the CPU oracle (C restatement + torch port) to the reference's own outputs (tests/golden/).

The reference has no tests of its own; these fixtures were produced by running its unmodified
functions in the build container (tests/golden/make_golden.py).
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port

REL = 1e-12  # fp64 oracle vs fp64 reference


def _close(a, b, rel=REL, abs_=1e-14):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300) if b.size else 1.0
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) <= rel * scale + abs_, (np.max(np.abs(a - b)), scale)


def _names(g):
    return [str(n) for n in g["names"]]


def test_energy_terms_match_reference(golden_energy):
    g = golden_energy
    for n in _names(g):
        beta = float(g[f"{n}/beta"])
        conf, inter, rc, ri = oracle.energy_terms(g[f"{n}/xhat"], g[f"{n}/x0"], beta)
        _close(conf, g[f"{n}/conf"])
        _close(inter, g[f"{n}/inter"])
        B, m, _ = g[f"{n}/xhat"].shape
        _close(rc.sum() / (B * m), g[f"{n}/conf"])
        _close(ri.sum() / (B * m * (m - 1)), g[f"{n}/inter"])


def test_energy_grads_match_reference_autograd(golden_energy):
    g = golden_energy
    for n in _names(g):
        beta = float(g[f"{n}/beta"])
        gc, gc0 = oracle.energy_terms_grad(g[f"{n}/xhat"], g[f"{n}/x0"], beta, 1.0, 0.0, want_x0=True)
        gi = oracle.energy_terms_grad(g[f"{n}/xhat"], g[f"{n}/x0"], beta, 0.0, 1.0)
        _close(gc, g[f"{n}/g_conf"], rel=1e-11)
        _close(gc0, g[f"{n}/g_conf_x0"], rel=1e-11)
        _close(gi, g[f"{n}/g_inter"], rel=1e-11)


def test_known_answers(golden_energy):
    """SURVEY.md §8(c) KAT table, through the oracle."""
    g = golden_energy
    conf, inter, *_ = oracle.energy_terms(g["kat1a/xhat"], g["kat1a/x0"], 2.0)
    assert (conf, inter) == (1.0, 4.0)
    loss, _, _, grad = oracle.energy_loss(g["kat1a/xhat"], g["kat1a/x0"], 2.0, 1.0, 1.0)
    assert loss == -1.0 and grad.ravel().tolist() == [-1.0, 1.0]
    loss, conf, inter, grad = oracle.energy_loss(g["kat1c/xhat"], g["kat1c/x0"], 0.1, 1.0, 1.0)
    assert abs(inter - 2 ** 0.1) < 1e-12 and abs(loss - 0.4641132687318966) < 1e-12
    assert np.allclose(np.abs(grad.ravel()), 0.023205663436551532, rtol=1e-10)
    loss, conf, inter, grad = oracle.energy_loss(g["kat2a/xhat"], g["kat2a/x0"], 1.0, 1.0, 1.0)
    assert abs(conf - 3.3333336666667) < 1e-9 and abs(inter - 6.66666666666675) < 1e-9
    assert np.allclose(grad.ravel(), [0.1, 0.4 / 3, 0, 0, -0.1, -0.4 / 3], atol=1e-7)
    loss, conf, inter, grad = oracle.energy_loss(g["kat3/xhat"], g["kat3/x0"], 0.1, 1.0, 1.0)
    assert abs(conf - 1e-12 ** 0.05) < 1e-12 and abs(inter - conf) < 1e-15 and not grad.any()
    assert np.all(g["lam_ignored"] == 0.0)


def test_sigmoid_weight(golden_weights):
    g = golden_weights
    for bias in (0.0, 1.0, -0.5):
        _close(oracle.sigmoid_weight(g["t"], bias), g[f"w64_bias{bias}"], rel=1e-13)
        w = torch_port.sigmoid_weight(torch.from_numpy(g["t"]), bias).numpy()
        _close(w, g[f"w64_bias{bias}"], rel=1e-13)
    w = oracle.sigmoid_weight(np.array([0.0, 0.25, 0.5, 0.75, 1.0]))
    assert np.allclose(w, [1, 0.9, 0.5, 0.1, 1e-12], rtol=1e-9)


def test_forward_marginal(golden_schedules):
    g = golden_schedules
    for name in ("img", "flat", "toy"):
        x0, eps, t = g[f"fm_{name}/x0"], g[f"fm_{name}/eps"], g[f"fm_{name}/t"]
        xt, rep = oracle.forward_marginal(x0, t, eps, m=3)
        _close(xt.reshape(x0.shape), g[f"fm_{name}/xt64"], rel=1e-15)
        assert np.array_equal(rep.reshape(x0.shape[0], 3, -1)[:, 1], xt)
        # the torch port run in fp32 reproduces the reference's fp32 output bit for bit
        xt32 = torch_port.forward_marginal(torch.from_numpy(x0), torch.from_numpy(t), torch.from_numpy(eps))
        assert np.array_equal(xt32.numpy(), g[f"fm_{name}/xt"])
    xt32 = torch_port.forward_marginal(*(torch.from_numpy(g[f"fm_lowrank/{k}"]) for k in ("x0", "t", "eps")))
    assert np.array_equal(xt32.numpy(), g["fm_lowrank/xt"])


def test_bridge(golden_schedules):
    g = golden_schedules
    x0hat, xt = g["br/x0hat"], g["br/xt"]
    for steps in (20, 5):
        grid32 = torch.linspace(0.0, 1.0, steps + 1)
        for churn in (1.0, 0.0, 0.5):
            key = f"br_grid{steps}_churn{churn}"
            for k in range(steps):
                s, t = float(grid32[k]), float(grid32[k + 1])
                _, mu = oracle.bridge_step(x0hat * 0 + xt, x0hat, None, s, t, churn)
                _close(mu, g[f"{key}/mu64"][k], rel=1e-13)
                assert abs(oracle.bridge_coeffs(s, t, churn)[2] - g[f"{key}/std64"][k]) < 1e-13
                mu32, std32 = torch_port.bridge(grid32[k], grid32[k + 1], torch.from_numpy(x0hat),
                                                torch.from_numpy(xt), churn)
                assert np.array_equal(mu32.numpy(), g[f"{key}/mu"][k])
                assert float(std32) == g[f"{key}/std"][k]
    # SURVEY §8(c) KAT 6
    c = oracle.bridge_coeffs(19 / 20, 1.0, 1.0)
    assert abs(c[0]) < 1e-7 and abs(c[1] - 0.05) < 1e-7 and abs(c[2] - 0.95) < 1e-7
    c = oracle.bridge_coeffs(0.5, 0.55, 1.0)
    assert np.allclose(c, (0.743802, 0.128099, 0.287480), atol=2e-6)
    assert oracle.bridge_coeffs(0.0, 0.05, 1.0) == (0.0, 1.0, 0.0)
    s, t = g["br_vec/s"], g["br_vec/t"]
    for churn in (1.0, 0.0, 0.7):
        nxt, mu = oracle.bridge_step(xt, x0hat, np.ones_like(xt), s, t, churn)
        assert np.allclose(mu, g[f"br_vec_churn{churn}/mu"], rtol=2e-6, atol=2e-6)  # golden is fp32
        assert np.allclose(nxt - mu, np.broadcast_to(g[f"br_vec_churn{churn}/std"], xt.shape), rtol=2e-6, atol=3e-7)
        mu32, std32 = torch_port.bridge(torch.from_numpy(s), torch.from_numpy(t), torch.from_numpy(x0hat),
                                        torch.from_numpy(xt), churn)
        assert np.array_equal(mu32.numpy(), g[f"br_vec_churn{churn}/mu"])
        assert np.array_equal(std32.numpy(), g[f"br_vec_churn{churn}/std"])
    mu32, std32 = torch_port.bridge(torch.from_numpy(s), torch.from_numpy(t), torch.from_numpy(g["br_img/x0hat"]),
                                    torch.from_numpy(g["br_img/xt"]), 1.0)
    assert np.array_equal(mu32.numpy(), g["br_img/mu"]) and std32.shape == (5, 1, 1, 1)


def test_training_step_loss_and_grad(golden_step):
    """training.py:77-85 on the recorded denoiser output: oracle (fp64) vs reference (fp32)."""
    g = golden_step
    for n in _names(g):
        m, beta, lam, w_bias = g[f"{n}/hyper"]
        m = int(m)
        xhat, x0, t = g[f"{n}/xhat"], g[f"{n}/x0"], g[f"{n}/t"]
        B = x0.shape[0]
        w = oracle.sigmoid_weight(t, w_bias).mean()
        loss, conf, inter, grad = oracle.energy_loss(xhat.reshape(B, m, -1), x0.reshape(B, -1), beta, lam, w)
        ref = g[f"{n}/scalars"]
        assert np.allclose([loss, conf, inter, w], ref, rtol=2e-6, atol=1e-7), (n, [loss, conf, inter, w], ref)
        gref = g[f"{n}/grad_xhat"].reshape(B, m, -1)
        assert np.max(np.abs(grad - gref)) <= 5e-6 * np.max(np.abs(gref)), n
        # forward marginal + expansion as the model saw them
        xt, _ = oracle.forward_marginal(x0, t, g[f"{n}/eps"])
        assert np.allclose(xt.reshape(x0.shape), g[f"{n}/xt"], rtol=1e-6, atol=1e-7)


def test_torch_port_step_bitwise(golden_step):
    """The torch port replays the reference's fp32 step bit for bit when given the same noise."""
    from tests.helpers import MixModel

    g = golden_step
    for n in _names(g):
        m, beta, lam, w_bias = g[f"{n}/hyper"]
        model = MixModel()
        loss, conf, inter, weight, xhat = torch_port.training_step(
            model, torch.from_numpy(g[f"{n}/x0"]), torch.from_numpy(g[f"{n}/t"]), torch.from_numpy(g[f"{n}/eps"]),
            torch.from_numpy(g[f"{n}/xi"]), m=int(m), beta=float(beta), lam=float(lam), w_bias=float(w_bias))
        xhat.retain_grad()
        loss.backward()
        assert np.array_equal(xhat.detach().numpy(), g[f"{n}/xhat"])
        got = np.array([float(v.detach()) for v in (loss, conf, inter, weight)])
        assert np.allclose(got, g[f"{n}/scalars"], rtol=1e-6), n
        assert np.allclose(xhat.grad.numpy(), g[f"{n}/grad_xhat"], rtol=1e-5, atol=1e-9)
    with pytest.raises(ValueError):
        torch_port.training_step(MixModel(), torch.zeros(2, 2), torch.zeros(2), torch.zeros(2, 2),
                                 torch.zeros(2, 1, 2), m=1, beta=1.0, lam=1.0, w_bias=0.0)
    assert "m must be >= 2" in str(g["m_lt_2_message"])


def test_sampler(golden_sampler):
    from tests.helpers import MixModel

    g = golden_sampler
    for n in _names(g):
        steps, churn = g[f"{n}/hyper"]
        steps = int(steps)
        x = torch_port.sample(MixModel().eval(), torch.from_numpy(g[f"{n}/x_init"]), torch.from_numpy(g[f"{n}/xis"]),
                              torch.from_numpy(g[f"{n}/zs"]), steps, float(churn))
        assert np.array_equal(x.numpy(), g[f"{n}/x_final"]), n
        # and the C oracle's update, chained over the same noise in fp64, stays within fp32 rounding
        model = MixModel().double()
        xx = g[f"{n}/x_init"].astype(np.float64)
        grid = torch.linspace(0.0, 1.0, steps + 1)
        for k in reversed(range(steps)):
            s, t = float(grid[k]), float(grid[k + 1])
            with torch.no_grad():
                xh = model(torch.from_numpy(xx), torch.full((xx.shape[0],), t, dtype=torch.float64),
                           torch.from_numpy(g[f"{n}/xis"][k].astype(np.float64))).numpy()
            xx, _ = oracle.bridge_step(xx, xh, g[f"{n}/zs"][k], s, t, churn)
        assert np.allclose(xx, g[f"{n}/x_final"], rtol=2e-5, atol=2e-5), n
    assert tuple(g["default_shape"]) == (3, 2)


def test_rbf_mmd2_oracle_matches_reference(golden_mmd):
    """oracle.rbf_mmd2 (numpy fp64 restatement of dddm/metrics.py:140-163) against the reference's own outputs."""
    g = golden_mmd
    for n in [str(s) for s in g["names"]]:
        mmd2, kxx, kyy, kxy = oracle.rbf_mmd2(g[f"{n}/x"], g[f"{n}/y"], float(g[f"{n}/sigma"]))
        assert abs(mmd2 - float(g[f"{n}/mmd2_f64"])) <= 1e-12 * max(kxx, kyy, kxy, 1e-300) + 1e-300, n
        # the reference's native fp32 run agrees to fp32 accuracy of the kernel means
        assert abs(float(g[f"{n}/mmd2_f32"]) - mmd2) <= 2e-6 * max(kxx, kyy, kxy, 1e-30) + 1e-30, n
    with pytest.raises(ValueError):
        oracle.rbf_mmd2(np.zeros((1, 2)), np.zeros((3, 2)))


def test_oracle_gradient_is_the_derivative_of_the_oracle_loss():
    """The C oracle's closed-form gradient against central finite differences of its own loss, and the
    invariances the GPU property tests rely on (common translation, permutation of the draws, permutation of D)."""
    rng = np.random.default_rng(0)
    for (B, m, D, beta) in ((2, 3, 5, 0.1), (1, 4, 7, 1.0), (2, 2, 3, 1.7), (1, 5, 4, 2.0)):
        x0 = rng.normal(size=(B, D))
        xh = x0[:, None, :] + 0.7 * rng.normal(size=(B, m, D))
        loss, conf, inter, grad = oracle.energy_loss(xh, x0, beta, 1.3, 0.6)
        num = np.zeros_like(xh)
        h = 1e-6
        for idx in np.ndindex(*xh.shape):
            xp, xm = xh.copy(), xh.copy()
            xp[idx] += h
            xm[idx] -= h
            num[idx] = (oracle.energy_loss(xp, x0, beta, 1.3, 0.6, want_grad=False)[0] -
                        oracle.energy_loss(xm, x0, beta, 1.3, 0.6, want_grad=False)[0]) / (2 * h)
        assert np.max(np.abs(num - grad)) <= 1e-7 * max(np.max(np.abs(grad)), 1e-12) + 1e-9, (B, m, D, beta)
        shift = rng.normal(size=(B, 1, D))
        l2, c2, i2, g2 = oracle.energy_loss(xh + shift, x0 + shift[:, 0], beta, 1.3, 0.6)
        assert abs(l2 - loss) <= 1e-12 * abs(loss) + 1e-14 and np.allclose(g2, grad, rtol=1e-9, atol=1e-14)
        perm, dperm = rng.permutation(m), rng.permutation(D)
        l3, _, _, g3 = oracle.energy_loss(xh[:, perm][:, :, dperm], x0[:, dperm], beta, 1.3, 0.6)
        assert abs(l3 - loss) <= 1e-12 * abs(loss) + 1e-14 and np.allclose(g3, grad[:, perm][:, :, dperm], rtol=1e-9, atol=1e-14)
