import os, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kin40k():
    return dict(np.load(os.path.join(GOLDEN, "kin40k_chain.npz")))


@pytest.fixture(scope="session")
def banana():
    return dict(np.load(os.path.join(GOLDEN, "banana_chain.npz")))


@pytest.fixture(scope="session")
def toy():
    return dict(np.load(os.path.join(GOLDEN, "toy_sets.npz")))
