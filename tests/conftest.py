import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_libraries():
    """Build the CUDA library (nvcc cross-compiles without a GPU) and the C oracle when they are
    missing or stale: the .so files are git-ignored, a fresh checkout has none."""
    from mpc_motion_planning_b200 import build
    from oracle import c_oracle

    build.build()
    c_oracle.build()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    path = os.path.join(ROOT, "tests", "golden", "solutions.npz")
    d = np.load(path)
    cases = {}
    for name in d["names"]:
        name = str(name)
        cases[name] = {k.split("/", 1)[1]: d[k] for k in d.files if k.startswith(name + "/")}
    return cases
