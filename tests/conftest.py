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
def pima():
    return dict(np.load(os.path.join(GOLDEN, "pima.npz")))


@pytest.fixture(scope="session")
def synth():
    return dict(np.load(os.path.join(GOLDEN, "synth2000x32.npz")))


def ungrad_scale(X, y, beta, pscale):
    """Un-cancelled magnitude of glp: |X|'|y-p| + |beta/pscale^2|. Relative
    gradient errors are measured against its max (hard part 3:
    glp -> 0 at the MAP, so a component-wise relative error is meaningless there)."""
    X = np.asarray(X, dtype=np.float64)
    pr = 1 / (1 + np.exp(-X.dot(beta)))
    return float(np.max(np.abs(X).T.dot(np.abs(y - pr)) + np.abs(beta / pscale ** 2)))
