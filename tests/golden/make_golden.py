#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE's own functions.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``dddm`` from ``/root/reference`` with a stub ``matplotlib`` in ``sys.modules``
(``dddm/data.py:9`` imports pyplot at module level and matplotlib is not installed), calls the
reference's unmodified ``generalized_energy_terms``, ``sigmoid_weight``, ``forward_marginal_sample``,
``gaussian_bridge_mu_sigma``, ``distributional_training_step`` and ``sample_dddm`` on seeded
inputs, and stores inputs + outputs as small ``.npz`` files.  The fixtures pin the CPU oracle
(``tests/test_oracle_golden.py``) and are compared with the CUDA kernels in the ``-m gpu`` tests.

Noise "passed in": the reference draws ``t, eps, xi`` (training.py:65-69) and ``x_T, xi_k, z_k``
(sampling.py:23-30) internally; we recover them by re-seeding and replaying the same draw order,
and verify the replay against what the model actually received.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("DDDM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import dddm  # noqa: F401
    from dddm import losses, sampling, schedules, training

    return losses, schedules, sampling, training


def f32_grid(t: torch.Tensor) -> torch.Tensor:
    """Round to fp32 and return as fp64 so both fp32 and fp64 consumers see identical values."""
    return t.float().double()


def energy_cases(losses):
    out = {}
    names = []

    def add(name, xhat, x0, beta):
        xh = xhat.double().clone().requires_grad_(True)
        c = x0.double().clone().requires_grad_(True)
        conf, inter = losses.generalized_energy_terms(xh, c, beta, 1.0)
        (g_conf, g_conf_x0) = torch.autograd.grad(conf, (xh, c), retain_graph=True)
        (g_inter,) = torch.autograd.grad(inter, (xh,), retain_graph=True)
        # the reference run natively in fp32 (its only shipped precision), for information
        conf32, inter32 = losses.generalized_energy_terms(xhat.float(), x0.float(), beta, 1.0)
        # inputs are fp32-representable, so storing them as fp32 is lossless
        for k, v in dict(xhat=xhat.float(), x0=x0.float(), beta=torch.tensor(beta, dtype=torch.float64),
                         conf=conf.detach(), inter=inter.detach(), g_conf=g_conf, g_inter=g_inter,
                         g_conf_x0=g_conf_x0, conf32=conf32.double(), inter32=inter32.double()).items():
            out[f"{name}/{k}"] = v.numpy()
        names.append(name)

    # known-answer cases of SURVEY.md §8(c)
    kat1 = (torch.tensor([[[1.0], [-1.0]]]), torch.tensor([[0.0]]))
    for tag, beta in (("a", 2.0), ("b", 1.0), ("c", 0.1)):
        add(f"kat1{tag}", *kat1, beta)
    kat2 = (torch.tensor([[[3.0, 4.0], [0.0, 0.0], [-3.0, -4.0]]]), torch.tensor([[0.0, 0.0]]))
    add("kat2a", *kat2, 1.0)
    add("kat2b", *kat2, 2.0)
    same = torch.tensor([[0.3, -0.7]])
    add("kat3", same[:, None, :].expand(1, 2, 2).clone(), same, 0.1)

    g = torch.Generator().manual_seed(20261018)
    shapes = [(3, 2, 1), (2, 3, 2), (4, 4, 5), (5, 8, 2), (3, 8, 33), (2, 16, 7), (2, 5, 130), (1, 32, 9),
              (2, 8, 259), (3, 7, 64)]
    for (B, m, D) in shapes:
        for beta in ((0.1, 1.0, 2.0, 1.5) if B * m * D <= 300 else (0.1, 2.0)):
            x0 = f32_grid(torch.randn(B, D, generator=g).clamp(-1, 1))
            early = f32_grid(torch.randn(B, m, D, generator=g))
            add(f"early_B{B}_m{m}_D{D}_beta{beta}", early, x0, beta)
            late = f32_grid(x0[:, None, :] + 0.05 * torch.randn(B, m, D, generator=g))
            add(f"late_B{B}_m{m}_D{D}_beta{beta}", late, x0, beta)
    # duplicate draws inside a random row (zero pair distance -> (1e-12)^(beta/2), zero gradient)
    for beta in (0.1, 1.0, 2.0):
        x0 = f32_grid(torch.randn(2, 6, generator=g))
        xh = f32_grid(torch.randn(2, 4, 6, generator=g))
        xh[0, 2] = xh[0, 0]
        xh[1, 1] = x0[1]
        add(f"dups_beta{beta}", xh, x0, beta)
    # bf16-representable inputs (bf16 kernels see exactly these values)
    for beta in (0.1, 2.0):
        x0 = torch.randn(3, 40, generator=g).clamp(-1, 1).bfloat16().double()
        xh = torch.randn(3, 8, 40, generator=g).bfloat16().double()
        add(f"bf16grid_beta{beta}", xh, x0, beta)
    # lam is ignored by generalized_energy_terms (losses.py:6): record two lam values
    xh, c = torch.randn(2, 3, 4, generator=g).double(), torch.randn(2, 4, generator=g).double()
    a = losses.generalized_energy_terms(xh, c, 1.0, 1.0)
    b = losses.generalized_energy_terms(xh, c, 1.0, 123.0)
    out["lam_ignored"] = np.array([float(a[0] - b[0]), float(a[1] - b[1])])
    out["names"] = np.array(names)
    return out


def weight_cases(losses):
    g = torch.Generator().manual_seed(7)
    t = torch.cat([torch.tensor([0.0, 0.25, 0.5, 0.75, 1.0, 1e-4, 1 - 1e-4]), torch.rand(25, generator=g)])
    out = {"t": t.double().numpy()}
    for bias in (0.0, 1.0, -0.5):
        out[f"w64_bias{bias}"] = losses.sigmoid_weight(t.double(), bias).numpy()
        out[f"w32_bias{bias}"] = losses.sigmoid_weight(t.float(), bias).double().numpy()
    return out


def schedule_cases(schedules):
    g = torch.Generator().manual_seed(11)
    out = {}
    # forward marginal, image-shaped and flat, fp32 as in the reference
    for name, shape in (("img", (4, 3, 2, 2)), ("flat", (6, 5)), ("toy", (8, 2))):
        x0 = torch.randn(*shape, generator=g)
        eps = torch.randn(*shape, generator=g)
        t = torch.rand(shape[0], generator=g)
        out[f"fm_{name}/x0"], out[f"fm_{name}/eps"], out[f"fm_{name}/t"] = x0.numpy(), eps.numpy(), t.numpy()
        out[f"fm_{name}/xt"] = schedules.forward_marginal_sample(x0, t, eps).numpy()
        out[f"fm_{name}/xt64"] = schedules.forward_marginal_sample(x0.double(), t.double(), eps.double()).numpy()
    # eps of lower rank broadcasts over trailing dims (schedules.py:20-21)
    x0 = torch.randn(3, 2, 4, generator=g)
    eps = torch.randn(3, 2, generator=g)
    t = torch.rand(3, generator=g)
    out["fm_lowrank/x0"], out["fm_lowrank/eps"], out["fm_lowrank/t"] = x0.numpy(), eps.numpy(), t.numpy()
    out["fm_lowrank/xt"] = schedules.forward_marginal_sample(x0, t, eps).numpy()

    # bridge: scalar (0-dim) s,t on the sampler's grids, and per-sample vectors
    x0hat = torch.randn(5, 3, generator=g)
    xt = torch.randn(5, 3, generator=g)
    out["br/x0hat"], out["br/xt"] = x0hat.numpy(), xt.numpy()
    for steps in (20, 5):
        grid = torch.linspace(0.0, 1.0, steps + 1)
        for churn in (1.0, 0.0, 0.5):
            mus, stds, mus64, stds64 = [], [], [], []
            for k in range(steps):
                mu, std = schedules.gaussian_bridge_mu_sigma(grid[k], grid[k + 1], x0hat, xt, eps_churn=churn)
                assert std.shape == (1, 1)
                mus.append(mu.numpy()), stds.append(float(std))
                mu, std = schedules.gaussian_bridge_mu_sigma(grid[k].double(), grid[k + 1].double(), x0hat.double(),
                                                             xt.double(), eps_churn=churn)
                mus64.append(mu.numpy()), stds64.append(float(std))
            out[f"br_grid{steps}_churn{churn}/mu"] = np.stack(mus)
            out[f"br_grid{steps}_churn{churn}/std"] = np.array(stds)
            out[f"br_grid{steps}_churn{churn}/mu64"] = np.stack(mus64)
            out[f"br_grid{steps}_churn{churn}/std64"] = np.array(stds64)
    s = torch.tensor([0.0, 0.1, 0.45, 0.5, 0.95])
    t = torch.tensor([0.05, 0.3, 0.5, 1.0, 1.0])
    out["br_vec/s"], out["br_vec/t"] = s.numpy(), t.numpy()
    for churn in (1.0, 0.0, 0.7):
        mu, std = schedules.gaussian_bridge_mu_sigma(s, t, x0hat, xt, eps_churn=churn)
        assert std.shape == (5, 1)
        out[f"br_vec_churn{churn}/mu"], out[f"br_vec_churn{churn}/std"] = mu.numpy(), std.numpy()
    # image-shaped vector case: std is right-padded to x0.ndim
    x4, h4 = torch.randn(5, 2, 3, 3, generator=g), torch.randn(5, 2, 3, 3, generator=g)
    mu, std = schedules.gaussian_bridge_mu_sigma(s, t, h4, x4, eps_churn=1.0)
    assert std.shape == (5, 1, 1, 1)
    out["br_img/x0hat"], out["br_img/xt"], out["br_img/mu"], out["br_img/std"] = (h4.numpy(), x4.numpy(), mu.numpy(),
                                                                              std.numpy())
    return out


class MixModel(torch.nn.Module):
    """Tiny deterministic stand-in for the denoiser: works for any trailing shape."""

    def __init__(self, a=0.8, b=0.35, c=-0.2):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor(a))
        self.b = torch.nn.Parameter(torch.tensor(b))
        self.c = torch.nn.Parameter(torch.tensor(c))
        self.seen = None
        self.out = None

    def forward(self, xt, t, xi):
        tt = t.reshape(t.shape + (1,) * (xt.ndim - 1))
        y = self.a * xt + self.b * xi * (1.0 + tt) + self.c * torch.tanh(xt * xi) + 0.1 * tt
        self.seen = (xt.detach().clone(), t.detach().clone(), xi.detach().clone())
        if y.requires_grad:
            y.retain_grad()
        self.out = y
        return y


def step_cases(training, schedules):
    out = {}
    names = []
    cfgs = [("toy", (16, 2), 8, 0.1, 1.0, 0.0, False), ("img", (4, 3, 4, 4), 4, 1.0, 0.7, 0.5, False),
            ("beta2", (5, 6), 3, 2.0, 1.3, -0.25, False), ("given_t", (6, 2, 3), 5, 1.5, 1.0, 0.0, True)]
    for idx, (name, shape, m, beta, lam, w_bias, give_t) in enumerate(cfgs):
        seed = 1000 + idx
        g = torch.Generator().manual_seed(seed)
        x0 = torch.randn(*shape, generator=g).clamp(-1, 1)
        t_in = torch.rand(shape[0], generator=g) if give_t else None
        model = MixModel()
        torch.manual_seed(seed)
        loss, metrics = training.distributional_training_step(model, x0, m=m, beta=beta, lam=lam, w_bias=w_bias,
                                                              t=t_in)
        loss.backward()
        # replay the RNG stream: rand(B) [unless t given] -> randn_like(x0) -> randn(B,m,...)  (training.py:65-69)
        torch.manual_seed(seed)
        t = t_in if give_t else torch.rand(shape[0])
        eps = torch.randn_like(x0)
        xi = torch.randn((shape[0], m, *shape[1:]))
        xt = schedules.forward_marginal_sample(x0, t, eps)
        seen_xt, seen_t, seen_xi = model.seen
        assert torch.equal(seen_xt.view(shape[0], m, *shape[1:])[:, 0], xt), "RNG replay mismatch (xt)"
        assert torch.equal(seen_xi, xi.reshape(shape[0] * m, *shape[1:])), "RNG replay mismatch (xi)"
        assert torch.equal(seen_t, t.repeat_interleave(m))
        rec = dict(x0=x0, t=t, eps=eps, xi=xi, xt=xt, xhat=model.out.detach().view(shape[0], m, *shape[1:]),
                   grad_xhat=model.out.grad.view(shape[0], m, *shape[1:]),
                   scalars=torch.tensor([metrics["loss"], metrics["confidence"], metrics["interaction"],
                                         metrics["weight"]], dtype=torch.float64),
                   hyper=torch.tensor([m, beta, lam, w_bias], dtype=torch.float64),
                   param_grads=torch.stack([model.a.grad, model.b.grad, model.c.grad]))
        for k, v in rec.items():
            out[f"{name}/{k}"] = v.numpy()
        names.append(name)
    out["names"] = np.array(names)
    # m < 2 must raise ValueError (training.py:57-58)
    try:
        training.distributional_training_step(MixModel(), torch.zeros(2, 2), m=1, beta=1.0, lam=1.0, w_bias=0.0)
        raised = ""
    except ValueError as e:
        raised = str(e)
    out["m_lt_2_message"] = np.array(raised)
    return out


def sampler_cases(sampling):
    out = {}
    names = []
    cfgs = [("toy_churn1", 6, (2,), 5, 1.0), ("toy_churn0", 6, (2,), 5, 0.0), ("img_churn03", 3, (3, 4, 4), 4, 0.3),
            ("toy_20", 4, (2,), 20, 1.0)]
    for idx, (name, n, shape, steps, churn) in enumerate(cfgs):
        seed = 2000 + idx
        model = MixModel()
        torch.manual_seed(seed)
        x = sampling.sample_dddm(model, n_samples=n, steps=steps, eps_churn=churn, device="cpu", data_shape=shape)
        assert not model.training
        # replay: randn(x_T); per step (k = steps-1 .. 0): randn_like (xi), randn_like (z)   (sampling.py:23-30)
        torch.manual_seed(seed)
        x_init = torch.randn((n, *shape))
        xis, zs = torch.empty(steps, n, *shape), torch.empty(steps, n, *shape)
        for k in reversed(range(steps)):
            xis[k] = torch.randn_like(x_init)
            zs[k] = torch.randn_like(x_init)
        rec = dict(x_init=x_init, xis=xis, zs=zs, x_final=x,
                   hyper=torch.tensor([steps, churn], dtype=torch.float64))
        for k, v in rec.items():
            out[f"{name}/{k}"] = v.numpy()
        names.append(name)
    out["names"] = np.array(names)
    # default data_shape is (2,) (sampling.py:21-22)
    torch.manual_seed(5)
    out["default_shape"] = np.array(sampling.sample_dddm(MixModel(), n_samples=3, steps=2).shape)
    return out


def mmd_cases():
    """rbf_mmd2 (dddm/metrics.py:140-163) on seeded sets: the toy GMM shape (D=2, sigma=1, run_example.py:101),
    a mid-dimensional case with a matched bandwidth, identical sets, and flattened images at sigma=1 (where every
    off-diagonal kernel value underflows to 0, as in compute_image_mmd's default)."""
    from dddm import metrics

    out, names = {}, []
    gen = torch.Generator().manual_seed(11)
    cases = {
        "toy": (torch.randn(600, 2, generator=gen) * 0.5 + torch.tensor([3.0, 3.0]),
                torch.randn(500, 2, generator=gen) * 0.6 + torch.tensor([2.7, 3.1]), 1.0),
        "mid": (torch.randn(300, 48, generator=gen), torch.randn(280, 48, generator=gen) * 1.1 + 0.1, 7.0),
        "same": (None, None, 2.0),
        "images": (torch.rand(12, 3072, generator=gen) * 2 - 1, torch.rand(10, 3072, generator=gen) * 2 - 1, 1.0),
        "images_wide": (torch.rand(24, 3072, generator=gen) * 2 - 1, torch.rand(20, 3072, generator=gen) * 1.8 - 0.9, 40.0),
    }
    z = torch.randn(200, 5, generator=gen)
    cases["same"] = (z, z.clone(), 2.0)
    for name, (x, y, sigma) in cases.items():
        v64 = metrics.rbf_mmd2(x.double(), y.double(), sigma)
        v32 = metrics.rbf_mmd2(x.float(), y.float(), sigma)
        out[f"{name}/x"], out[f"{name}/y"] = x.float().numpy(), y.float().numpy()
        out[f"{name}/sigma"] = np.float64(sigma)
        out[f"{name}/mmd2_f64"], out[f"{name}/mmd2_f32"] = v64.numpy(), v32.numpy()
        names.append(name)
    out["names"] = np.array(names)
    return out


def main():
    torch.set_num_threads(1)  # deterministic reductions for the recorded fp32 numbers
    losses, schedules, sampling, training = import_reference()
    np.savez_compressed(os.path.join(HERE, "energy.npz"), **energy_cases(losses))
    np.savez_compressed(os.path.join(HERE, "weights.npz"), **weight_cases(losses))
    np.savez_compressed(os.path.join(HERE, "schedules.npz"), **schedule_cases(schedules))
    np.savez_compressed(os.path.join(HERE, "step.npz"), **step_cases(training, schedules))
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **sampler_cases(sampling))
    np.savez_compressed(os.path.join(HERE, "mmd.npz"), **mmd_cases())
    meta = f"torch {torch.__version__}, numpy {np.__version__}, reference at {REF}\n"
    with open(os.path.join(HERE, "PROVENANCE.txt"), "w") as f:
        f.write("Generated by tests/golden/make_golden.py from the unmodified reference functions.\n" + meta)
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
