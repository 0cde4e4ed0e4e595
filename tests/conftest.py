import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def rubber_whale():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "rubber_whale_u8.npz")))


@pytest.fixture(scope="session")
def notebook_runs():
    import json
    with open(os.path.join(GOLDEN, "notebook_trajectories.json")) as f:
        return json.load(f)["notebooks"]


@pytest.fixture(scope="session")
def reference_runs():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "reference_runs.npz")))
