import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rms(x) -> float:
    return float(np.sqrt(np.mean(np.square(np.asarray(x, dtype=np.float64)))))


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
        return cache[name]
    return load


@pytest.fixture(scope="session")
def canonical():
    """preset name -> canonical folded weights of the seed-0 synthetic checkpoint."""
    from fastenhancer_b200.config import PRESETS
    from fastenhancer_b200.fold import fold_to_canonical
    from fastenhancer_b200.schema import synthetic_state_dict
    cache = {}

    def get(name, seed=0):
        if (name, seed) not in cache:
            cache[(name, seed)] = fold_to_canonical(PRESETS[name], synthetic_state_dict(PRESETS[name], seed))
        return cache[(name, seed)]
    return get
