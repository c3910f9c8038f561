import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "standin_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_samples():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "standin_samples.npz")))
