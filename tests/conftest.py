import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "n" in d:
        idx = np.cumsum(d["n"])[:-1]
        d["tb"], d["yb"], d["sb"] = np.split(d["t"], idx), np.split(d["y"], idx), np.split(d["s"], idx)
    if "kernel" in d:
        d["kernel"] = str(d["kernel"])
    return d


@pytest.fixture(scope="session")
def ctx():
    """The library context on cuda:0.  GPU tests must fail (not skip) when the CUDA path is unavailable."""
    import gpcc_b200
    return gpcc_b200.Context(1, profiling=True)
