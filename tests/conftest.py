"""Shared test plumbing.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g

    g.build()
    return g


@pytest.fixture(scope="session")
def orc(built):
    """The CPU oracle (checker only)."""
    import oracle

    return oracle


@pytest.fixture(scope="session")
def ck(built):
    """The product package (ctypes over libckks_b200.so)."""
    return importlib.import_module("toy-heaan-ckks_b200")


@pytest.fixture(scope="session")
def gpu(ck):
    if ck.device_count() < 1:
        pytest.fail("gpu-marked test started without a CUDA device (no CPU fallback exists)")
    return ck


def uniform_limbs(rng, moduli, n, *lead):
    """Words uniform in [0, q_limb), shape [*lead, L, n]."""
    q = np.array(moduli, dtype=np.uint64)
    raw = rng.integers(0, 1 << 63, size=(*lead, len(moduli), n), dtype=np.uint64)
    return (raw % q[:, None]).astype(np.uint64)
