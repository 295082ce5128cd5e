"""Shared pytest configuration.

* ``-m "not gpu"`` (runs in the CPU-only build container): oracle vs golden vectors, host logic,
  C-ABI symbol checks, gloo world_size-2 tests.
* ``-m gpu`` (runs on a B200): parity of the CUDA kernels, called through the C-ABI, against the
  CPU oracle and the golden fixtures.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_energy():
    return load_golden("energy.npz")


@pytest.fixture(scope="session")
def golden_weights():
    return load_golden("weights.npz")


@pytest.fixture(scope="session")
def golden_schedules():
    return load_golden("schedules.npz")


@pytest.fixture(scope="session")
def golden_step():
    return load_golden("step.npz")


@pytest.fixture(scope="session")
def golden_sampler():
    return load_golden("sampler.npz")


@pytest.fixture(scope="session")
def golden_mmd():
    return load_golden("mmd.npz")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
