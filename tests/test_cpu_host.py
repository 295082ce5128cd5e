"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol,
argument validation, the CUDA-only guards, the reference patching and the metrics mapping.
No compute calls (there is no GPU in the build container)."""
import ctypes
import os
import re
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cabi():
    from ddm_b200 import _cabi, build

    if not os.path.exists(_cabi.LIB_PATH):
        build.build()
    return _cabi


def test_library_exports_every_declared_symbol(cabi):
    header = open(os.path.join(ROOT, "include", "dddm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dddm_[a-z0-9_]+)\s*\(", header))
    declared -= {"dddm_session"}  # the opaque struct
    assert len(declared) >= 25
    L = ctypes.CDLL(cabi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/dddm_b200.h but not exported"
    assert declared == set(cabi.SIGNATURES), declared ^ set(cabi.SIGNATURES)
    assert cabi.lib().dddm_abi_version() == 1


def test_status_messages_and_validation(cabi):
    L = cabi.lib()
    assert cabi.strerror(0) == "ok"
    assert "m >= 2" in cabi.strerror(-2)
    with pytest.raises(ValueError):
        cabi.check(-2)
    with pytest.raises(cabi.DDDMError):
        cabi.check(-1)
    # argument validation happens before anything touches the device
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    assert L.dddm_energy_fused_f32(None, p, p, 1.0, None, p, p, 4, 8, 16, 0.1, 1.0, None) == -1
    assert L.dddm_energy_fused_f32(p, p, p, 1.0, None, p, p, 4, 1, 16, 0.1, 1.0, None) == -2  # m < 2
    assert L.dddm_energy_fused_f32(p, p, p, 1.0, None, p, p, 0, 8, 16, 0.1, 1.0, None) == -2
    assert L.dddm_energy_fused_f32(p, p, p, 1.0, None, p, p + 4, 4, 8, 16, 0.1, 1.0, None) == -3  # workspace alignment
    assert L.dddm_energy_terms_bwd_f32(p, p, None, p, p, p, None, 4, 8, 16, 0.1, None) == -1
    assert L.dddm_forward_marginal_expand_f32(p, p, p, None, None, 4, 8, 16, None) == -1
    assert L.dddm_bridge_step_f32(None, p, p, p, p, p, 0, 1.0, None, None, 4, 4, None) == -1
    assert L.dddm_set_tuning(b"no.such.key", 1) == -5
    assert L.dddm_energy_workspace_bytes(128, 8) == 16 + 128 * 8
    assert L.dddm_energy_dist_per_row(8) == 36 and L.dddm_energy_dist_per_row(32) == 528


def test_kernel_plans(cabi):
    try:
        assert cabi.describe_energy(128, 8, 3072).startswith("smem<f32,M=8> tma-bulk f32x2")
        assert cabi.describe_energy(128, 8, 3072, "bf16").startswith("smem<bf16,M=8>")
        # small minibatches are split along D so that B x cluster CTAs cover the SMs (one CTA per SM at most)
        assert "cluster=1 " in cabi.describe_energy(128, 8, 3072) and "cluster=1 " in cabi.describe_energy(96, 8, 3072)
        assert "cluster=2 threads=128 slab_vecs=384" in cabi.describe_energy(64, 8, 3072)
        assert "cluster=4 threads=128 slab_vecs=192" in cabi.describe_energy(32, 8, 3072)
        assert "cluster=4 threads=96 slab_vecs=96" in cabi.describe_energy(16, 8, 3072, "bf16")
        assert "cluster=1 " in cabi.describe_energy(16, 8, 512)  # narrow rows stay whole
        cabi.set_tuning("energy.loader", 2)
        assert cabi.describe_energy(128, 8, 3072).startswith("smem<f32,M=8> cp.async f32x2")
        cabi.set_tuning("energy.loader", 0)
        cabi.set_tuning("energy.variant", 5)  # the single-wave register-resident kernel is opt-in
        assert cabi.describe_energy(128, 8, 3072).startswith("wave<f32,M=8,NV=3> ldg.128 f32x2 register-resident")
        assert "threads=256" in cabi.describe_energy(128, 8, 3072)
        assert cabi.describe_energy(128, 8, 12288) == "unsupported"  # row too wide for the register file
        cabi.set_tuning("energy.variant", 0)
        assert cabi.describe_energy(512, 8, 2).startswith("reg<f32,M=8,VEC=1")
        cabi.set_tuning("energy.variant", 1)
        assert cabi.describe_energy(128, 8, 3072).startswith("reg<f32,M=8,VEC=4")
        cabi.set_tuning("energy.variant", 0)
        assert cabi.describe_energy(128, 32, 3072).startswith("blk<f32,M=32> tma-bulk")
        # bf16 draws at m = 16 / 32 with D a multiple of 128: the tensor-core kernel; the blocked kernel otherwise / on request
        assert cabi.describe_energy(128, 16, 3072, "bf16").startswith("tc<bf16,M=16> tcgen05")
        assert cabi.describe_energy(128, 32, 3072, "bf16").startswith("tc<bf16,M=32> tcgen05")
        assert cabi.describe_energy(128, 32, 12288, "bf16").startswith("tc<bf16,M=32> tcgen05")
        assert cabi.describe_energy(128, 16, 3080, "bf16").startswith("blk<bf16,M=16> tma-bulk")
        cabi.set_tuning("energy.variant", 4)
        assert cabi.describe_energy(128, 16, 3072, "bf16").startswith("blk<bf16,M=16> tma-bulk")
        cabi.set_tuning("energy.variant", 6)
        assert cabi.describe_energy(128, 8, 3072).startswith("pipe<f32,M=8> tma-bulk f32x2 rows-per-cluster=4")
        cabi.set_tuning("energy.variant", 0)
        assert cabi.describe_energy(128, 24, 3072).startswith("blk<f32,M=24> tma-bulk")
        assert cabi.describe_energy(128, 20, 3072).startswith("tile<f32,m=20>") and "tma-bulk" in cabi.describe_energy(
            128, 20, 3072)
        assert "ldg" in cabi.describe_energy(4, 16, 7)
        assert cabi.describe_energy(4, 65, 8) == "unsupported"
        cabi.set_tuning("energy.variant", 2)
        assert cabi.describe_energy(128, 8, 3072).startswith("tile<f32,m=8>")
        cabi.set_tuning("energy.variant", 1)
        cabi.set_tuning("energy.cluster", 4)
        assert "cluster=4 threads=192" in cabi.describe_energy(128, 8, 3072)
    finally:
        cabi.set_tuning("energy.variant", 0)
        cabi.set_tuning("energy.cluster", 0)
        cabi.set_tuning("energy.loader", 0)


def test_cuda_only_guards():
    import ddm_b200

    with pytest.raises(RuntimeError, match="CUDA-only"):
        ddm_b200.sigmoid_weight(torch.rand(4))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        ddm_b200.forward_marginal_sample(torch.zeros(2, 3), torch.zeros(2), torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        ddm_b200.sample_dddm(torch.nn.Identity(), n_samples=2, steps=1, device="cpu")
    with pytest.raises(ValueError, match="m must be >= 2"):
        ddm_b200.distributional_training_step(torch.nn.Identity(), torch.zeros(2, 2), m=1, beta=1.0, lam=1.0,
                                              w_bias=0.0)
    a, s = ddm_b200.alpha_sigma(torch.tensor([0.25]))
    assert float(a) == 0.75 and float(s) == 0.25


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ddm_b200 import _cabi

    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.DDDMLibraryMissing, match="no CPU / eager fallback"):
        _cabi.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ddm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_deferred_metrics_mapping():
    from ddm_b200 import DeferredMetrics

    mtr = DeferredMetrics(torch.tensor([1.0, 2.0, 3.0, 0.5]))
    assert mtr._host is None
    assert list(mtr) == ["loss", "confidence", "interaction", "weight"] and len(mtr) == 4
    assert mtr["confidence"] == 2.0 and dict(mtr) == {"loss": 1.0, "confidence": 2.0, "interaction": 3.0, "weight": 0.5}
    assert all(isinstance(v, float) for v in mtr.values())


def test_train_config_defaults():
    from ddm_b200 import TrainConfig

    c = TrainConfig()
    assert (c.beta, c.lam, c.m, c.w_bias, c.lr, c.epochs, c.batch, c.device, c.seed) == (0.1, 1.0, 8, 0.0, 2e-3, 2000,
                                                                                       512, "cpu", 0)


def test_patch_reference_rebinds_imported_names():
    """The reference binds names with `from .losses import ...` (training.py:10,12; sampling.py:5)."""
    import sys

    import ddm_b200

    pkg = types.ModuleType("fake_dddm")
    subs = {}
    for name in ("losses", "schedules", "training", "sampling", "metrics"):
        mod = types.ModuleType(f"fake_dddm.{name}")
        subs[name] = mod
        setattr(pkg, name, mod)
        sys.modules[f"fake_dddm.{name}"] = mod
    sys.modules["fake_dddm"] = pkg
    sentinel = object()
    for mod, names in ((subs["losses"], ["generalized_energy_terms", "sigmoid_weight"]),
                       (subs["schedules"], ["forward_marginal_sample", "gaussian_bridge_mu_sigma"]),
                       (subs["training"], ["generalized_energy_terms", "sigmoid_weight", "forward_marginal_sample",
                                           "distributional_training_step"]),
                       (subs["sampling"], ["gaussian_bridge_mu_sigma", "sample_dddm"]),
                       (subs["metrics"], ["rbf_mmd2"])):
        for n in names:
            setattr(mod, n, sentinel)
    try:
        saved = ddm_b200.patch_reference(pkg)
        assert subs["training"].generalized_energy_terms is ddm_b200.generalized_energy_terms
        assert subs["training"].forward_marginal_sample is ddm_b200.forward_marginal_sample
        assert subs["training"].distributional_training_step is ddm_b200.distributional_training_step
        assert subs["sampling"].gaussian_bridge_mu_sigma is ddm_b200.gaussian_bridge_mu_sigma
        assert pkg.sample_dddm is ddm_b200.sample_dddm
        assert subs["metrics"].rbf_mmd2 is ddm_b200.rbf_mmd2 and pkg.rbf_mmd2 is ddm_b200.rbf_mmd2
        ddm_b200.unpatch_reference(saved)
        assert subs["metrics"].rbf_mmd2 is sentinel
        assert subs["training"].sigmoid_weight is sentinel and not hasattr(pkg, "sample_dddm")
    finally:
        for k in [k for k in sys.modules if k.startswith("fake_dddm")]:
            del sys.modules[k]


def test_sass_shows_tma_bulk_copies_packed_fp32_and_tcgen05():
    """What the built library actually contains (B200_PROFILING.md, 'What proves a Blackwell-native kernel'):
    the energy kernels stage their tiles with TMA bulk copies (SASS UBLKCP, completion on mbarriers: SYNCS) and do
    their arithmetic in packed fp32 (FFMA2 / FADD2); no legacy tensor-core path (HMMA) is linked in."""
    import shutil
    import subprocess

    from ddm_b200 import _cabi

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(_cabi.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run([cuobjdump, "-sass", _cabi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in sass
    per_kernel = sass.split("Function : ")
    smem = [k for k in per_kernel if "energy_fused_smem_kernelIfLi8" in k.split("\n", 1)[0]]
    blk = [k for k in per_kernel if "energy_fused_blk_kernelIfLi32" in k.split("\n", 1)[0]]
    assert smem and blk
    ldgsts = [k for k in smem if "ELb0ELi1ELb0ELi4EEEv" in k.split("\n", 1)[0]]  # the cp.async loader instantiation (LOADER = 1)
    assert ldgsts and len(ldgsts) < len(smem)
    for k in ldgsts:
        assert "LDGSTS" in k and "LDGDEPBAR" in k, "cp.async loader: LDGSTS + commit groups expected"
    for k in smem + blk:
        if k in ldgsts:
            continue
        assert "UBLKCP" in k and "SYNCS" in k, "TMA bulk copy + mbarrier expected"
    for k in smem + blk:
        assert k.count("FFMA2") > 80 and k.count("FADD2") > 80, "packed fp32 arithmetic expected"  # bwd-only build: 112 / 86
    # tensor cores: only through tcgen05 (SASS UTCHMMA, operands by TMA tensor loads UTMALDG, accumulators read from
    # tensor memory LDTM) — the m = 16 / 32 energy kernel and rbf_mmd2; no legacy mma.sync (HMMA.*) / wgmma (HGMMA) path
    import re

    assert not re.search(r"(?<!UTC)HMMA", sass) and "HGMMA" not in sass
    tc = [k for k in per_kernel if "energy_tc_kernelILi32" in k.split("\n", 1)[0]]
    mmd = [k for k in per_kernel if "rbf_gram_sum_tc_kernel" in k.split("\n", 1)[0]]
    assert tc and mmd
    for k in tc + mmd:
        assert k.count("UTCHMMA") >= 8 and "UTMALDG" in k and "LDTM" in k and "UTCBAR" in k, "tcgen05 + TMA + TMEM expected"
    assert "UTMASTG" in tc[0], "the gradient leaves through TMA tensor stores"
