"""Backbones (out of hot-path scope, PyTorch): architecture parity with the reference where it is importable."""
import os
import sys
import types

import pytest
import torch

from ddm_b200.backbones import DDDMDiT, DDDMMLP

REF = "/root/reference"


def test_parameter_counts_and_shapes():
    dit = DDDMDiT()
    assert sum(p.numel() for p in dit.parameters()) == 14_523_312
    x = torch.randn(2, 3, 32, 32)
    assert dit(x, torch.rand(2), torch.randn_like(x)).shape == (2, 3, 32, 32)
    mlp = DDDMMLP()
    assert mlp(torch.randn(5, 2), torch.rand(5), torch.randn(5, 2)).shape == (5, 2)
    with pytest.raises(ValueError):
        dit(x, torch.rand(2), torch.randn(2, 3, 16, 16))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_state_dict_and_forward_match_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    try:
        from dddm import model as ref
    finally:
        sys.path.remove(REF)
    torch.manual_seed(0)
    for ours, theirs, shape in ((DDDMDiT(depth=2), ref.DDDMDiT(depth=2), (3, 3, 32, 32)), (DDDMMLP(), ref.DDDMMLP(), (7, 2))):
        sd = theirs.state_dict()
        assert {k: v.shape for k, v in sd.items()} == {k: v.shape for k, v in ours.state_dict().items()}
        ours.load_state_dict(sd)
        xt, xi, t = torch.randn(*shape), torch.randn(*shape), torch.rand(shape[0])
        with torch.no_grad():
            assert torch.allclose(ours(xt, t, xi), theirs(xt, t, xi), rtol=1e-4, atol=1e-5)
