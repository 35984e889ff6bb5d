import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def golden_3d():
    return load_golden("golden_3d.json")


@pytest.fixture(scope="session")
def golden_axisym():
    return load_golden("golden_axisym.json")


@pytest.fixture(scope="session")
def golden_slice():
    return load_golden("golden_slice.json")


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
