import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_match():
    return np.load(os.path.join(GOLDEN, "match.npz"))


@pytest.fixture(scope="session")
def golden_ess():
    return np.load(os.path.join(GOLDEN, "essential.npz"))


@pytest.fixture(scope="session")
def ctx():
    """The product path: fails loudly if the CUDA library or device is missing."""
    from epivo_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def esame(Ea, Eb):
    """distance between two essential matrices up to scale and sign"""
    Ea = np.asarray(Ea, dtype=np.float64).reshape(3, 3)
    Eb = np.asarray(Eb, dtype=np.float64).reshape(3, 3)
    Ea = Ea / np.linalg.norm(Ea)
    Eb = Eb / np.linalg.norm(Eb)
    return min(np.abs(Ea - Eb).max(), np.abs(Ea + Eb).max())
