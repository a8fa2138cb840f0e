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


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def bench_blobs(n, d, centers, seed, dtype):
    """The synthetic generator of SURVEY.md section 8(d) (also tests/golden/make_golden.py)."""
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(centers, d)) * (30.0 / np.sqrt(d))
    lab = rng.integers(0, centers, size=n)
    return (c[lab] + rng.normal(size=(n, d))).astype(dtype)


def golden_strings():
    g = load_golden("strings")
    X = np.array(g["text"].tobytes().decode("ascii").split("\n"))
    return X, g


@pytest.fixture(scope="session")
def gpu_ctx():
    import annchor_b200 as ab
    return ab.default_context()
