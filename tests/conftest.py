import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU and the built libgpfq.so")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine; GPU tests FAIL (not skip) when it cannot be created -- there is no fallback."""
    from quantized_neural_networks_b200 import get_engine
    return get_engine(0)


def hidden_pair(rng, N0, m, noise=0.05):
    Z = rng.standard_normal((N0, m))
    X = np.maximum(Z, 0).astype(np.float32)
    Xq = np.maximum(Z + noise * rng.standard_normal((N0, m)), 0).astype(np.float32)
    return X, Xq


def glorot(rng, N0, N1):
    return (rng.uniform(-1, 1, (N0, N1)) * np.sqrt(6.0 / (N0 + N1))).astype(np.float32)
