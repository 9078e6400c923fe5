import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def signal(seed, rows, n, fs):
    """Same recipe as oracle/make_golden.py:signal."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    x = rng.standard_normal((rows, n))
    x += 20 * np.sin(2 * np.pi * 60 * t) + 30 * np.sin(2 * np.pi * 8 * t) + 5 * t / t[-1]
    return x


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture
def fake_gpu(monkeypatch):
    """Route the kernel-level calls of openseize_b200.core.device to numpy
    stand-ins that follow the C-ABI contracts, so the HOST logic (chunk
    bookkeeping, carries, laziness, device chaining) is testable without a
    GPU.  Test infrastructure only; the product has no such path."""
    from tests import fake_backend

    fake_backend.install(monkeypatch)
    yield


def has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
