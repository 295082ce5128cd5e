"""Helpers shared by the tests (not part of the product)."""
import torch


class MixModel(torch.nn.Module):
    """Tiny deterministic stand-in for the denoiser x0hat = f(xt, t, xi); any trailing shape.

    Same arithmetic as the ``MixModel`` used by tests/golden/make_golden.py to record the
    training-step and sampler fixtures, so replays are bit-reproducible on CPU.
    """

    def __init__(self, a=0.8, b=0.35, c=-0.2):
        super().__init__()
        self.a = torch.nn.Parameter(torch.tensor(a))
        self.b = torch.nn.Parameter(torch.tensor(b))
        self.c = torch.nn.Parameter(torch.tensor(c))

    def forward(self, xt, t, xi):
        tt = t.reshape(t.shape + (1,) * (xt.ndim - 1))
        return self.a * xt + self.b * xi * (1.0 + tt) + self.c * torch.tanh(xt * xi) + 0.1 * tt
