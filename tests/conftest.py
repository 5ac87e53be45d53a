import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than ~30 s on CPU")


def load_golden(name):
    """tests/golden/<name>.npz -> {case: {key: ndarray}}."""
    import numpy as np

    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False)
    cases = {}
    for k in z.files:
        c, f = k.split("/", 1)
        cases.setdefault(c, {})[f] = z[k]
    return cases


@pytest.fixture(scope="session")
def golden_medium():
    return load_golden("medium")


@pytest.fixture(scope="session")
def golden_appendix_b():
    return load_golden("appendix_b")


@pytest.fixture(scope="session")
def golden_canonical():
    return load_golden("canonical")
