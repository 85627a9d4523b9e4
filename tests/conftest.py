import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def shipped_weights():
    w = np.fromfile(os.path.join(GOLDEN, "weights.bin"), dtype=np.uint8)
    assert w.size == 23184
    return w


@pytest.fixture(scope="session")
def conv_golden():
    return np.load(os.path.join(GOLDEN, "conv_cases.npz"))


@pytest.fixture(scope="session")
def tail_golden():
    return np.load(os.path.join(GOLDEN, "tail_cases.npz"))


@pytest.fixture(scope="session")
def cam_golden():
    return np.load(os.path.join(GOLDEN, "cam_cases.npz"))


@pytest.fixture(scope="session")
def prep_golden():
    return np.load(os.path.join(GOLDEN, "prep_cases.npz"))


@pytest.fixture(scope="session")
def acc24_golden():
    return np.load(os.path.join(GOLDEN, "acc24_cases.npz"))


@pytest.fixture(scope="session")
def pil_golden():
    return np.load(os.path.join(GOLDEN, "pil_cases.npz"))


@pytest.fixture(scope="session")
def trainer_golden():
    return np.load(os.path.join(GOLDEN, "trainer_case.npz"))
