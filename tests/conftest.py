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


@pytest.fixture(scope="session")
def golden_liif():
    return np.load(os.path.join(GOLDEN, "liif.npz"))


# LIIF-proper cases shared by the golden generator, the oracle pin and the GPU parity tests:
# name -> (B, H, W, Q or (H_up, W_up), weight seed, gain, feature seed)   (tests/golden/make_golden_liif.py CASES)
LIIF_CASES = {
    "sampled": (2, 12, 17, 700, 5, 1.0, 21),
    "sampled_gain": (1, 9, 8, 500, 6, 4.0, 22),
    "grid_x3": (1, 10, 13, (30, 39), 5, 1.0, 23),
    "grid_odd": (2, 7, 9, (16, 25), 7, 2.0, 24),
}


def liif_case(golden, name):
    """-> weights, feat, coord (B,Q,2), cell (B,Q,2), {ens: reference output}"""
    from diinn_b200 import synth
    B, H, W, q, wseed, gain, fseed = LIIF_CASES[name]
    weights = synth.make_liif_weights(wseed, gain)
    feat = synth.make_feat(fseed, B, H, W)
    coord, cell = golden[f"{name}.coord"], golden[f"{name}.cell"]
    if coord.ndim == 2:   # grid cases store one image's coordinates
        coord = np.ascontiguousarray(np.broadcast_to(coord, (B,) + coord.shape))
        cell = np.ascontiguousarray(np.broadcast_to(cell, (B,) + cell.shape))
    return weights, feat, coord, cell, {e: golden[f"{name}.ens{e}"] for e in (0, 1)}
