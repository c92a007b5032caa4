import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def primate_genome():
    """primate.p (12 x 898) encoded by the reference's own loader lines (tests/golden/make_golden.py)."""
    import numpy as np
    z = np.load(os.path.join(GOLDEN, "loader.npz"))
    return z["primate_genome"].astype(np.float64)
