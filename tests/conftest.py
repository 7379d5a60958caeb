import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def qbm():
    """The product package, with the CUDA extension loaded (fails loudly when it is not built)."""
    import qbm_b200
    qbm_b200._lib.load()
    return qbm_b200


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def random_qubo(n, seed, density=1.0, scale=1.0):
    import numpy as np
    rng = np.random.default_rng(seed)
    Q = np.triu(rng.uniform(-1.0, 1.0, (n, n))) * scale
    if density < 1.0:
        mask = np.triu(rng.random((n, n)) < density, k=1) | np.eye(n, dtype=bool)
        Q = Q * mask
    return Q
