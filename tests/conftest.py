import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(ROOT / "tests" / "golden" / "golden.npz"))


@pytest.fixture(scope="session")
def example_obs(golden):
    return golden["example_obs"].astype(np.int8)
