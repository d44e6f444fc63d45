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
    config.addinivalue_line("markers", "slow: long CPU test")


@pytest.fixture(scope="session")
def gold():
    return np.load(os.path.join(GOLDEN, "simulation_log_golden.npz"))


@pytest.fixture(scope="session")
def gait_gold():
    return np.load(os.path.join(GOLDEN, "gait_golden.npz"))


@pytest.fixture(scope="session")
def swing_gold():
    return np.load(os.path.join(GOLDEN, "swing_golden.npz"))
