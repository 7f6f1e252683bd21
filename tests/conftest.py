import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
MESHES = ["mesh2_1", "mesh5_1", "mesh_fine_1"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session", params=MESHES)
def ops(request):
    return load_golden(request.param + "_ops")


@pytest.fixture(scope="session")
def ops5():
    return load_golden("mesh5_1_ops")
