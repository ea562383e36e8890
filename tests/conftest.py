import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: larger cases (still part of the gpu suite)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand from oracle/."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def prims_golden():
    return np.load(os.path.join(GOLDEN, "prims.npz"))


@pytest.fixture(scope="session")
def extract_golden():
    return np.load(os.path.join(GOLDEN, "extract.npz"))


@pytest.fixture(scope="session")
def fbe():
    """The CUDA library; on a GPU box a missing library is an ERROR, never a skip or a fallback."""
    from fishbirdeyevisualslam_b200 import _lib
    return _lib.load()


def rand_desc(rng, n):
    return rng.integers(0, 256, (n, 32), dtype=np.uint8)


def make_kps(x, y, octave=None, angle=None, response=None):
    from fishbirdeyevisualslam_b200._lib import KP_DTYPE
    n = len(x)
    k = np.zeros(n, KP_DTYPE)
    k["x"], k["y"] = x, y
    k["size"] = 31
    k["octave"] = 0 if octave is None else octave
    k["angle"] = 0 if angle is None else angle
    k["response"] = 20 if response is None else response
    k["class_id"] = -1
    return k
