import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "kwinto-cuda_b200"))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/fd1d_oracle.c).  Test infrastructure only."""
    import pyoracle

    pyoracle.build(ref=True)
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def reflib():
    import pyoracle

    if not pyoracle.RefLib.available():
        pytest.skip("oracle/_ref/libkwref.so not built (reference tree not mounted)")
    return pyoracle.RefLib()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def synthetic_cases():
    g = load_golden("synthetic")
    keys = sorted({k.split("/")[0] for k in g.files})
    return g, keys
