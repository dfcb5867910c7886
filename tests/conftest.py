import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


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
