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
def golden_posenc():
    return np.load(os.path.join(GOLDEN, "posenc.npz"))


@pytest.fixture(scope="session")
def golden_decoder():
    return np.load(os.path.join(GOLDEN, "decoder.npz"))


@pytest.fixture(scope="session")
def golden_eval():
    return np.load(os.path.join(GOLDEN, "eval.npz"))


@pytest.fixture(scope="session")
def golden_modes():
    return np.load(os.path.join(GOLDEN, "modes.npz"))


@pytest.fixture(scope="session")
def golden_mode4():
    return np.load(os.path.join(GOLDEN, "mode4.npz"))


@pytest.fixture(scope="session")
def golden_initq():
    return np.load(os.path.join(GOLDEN, "initq.npz"))
