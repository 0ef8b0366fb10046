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
        return json.load(f)


@pytest.fixture(scope="session")
def golden_points():
    return load_golden("points.json")


@pytest.fixture(scope="session")
def golden_msm():
    return load_golden("msm.json")


@pytest.fixture(scope="session")
def golden_pairing():
    return load_golden("pairing.json")


@pytest.fixture(scope="session")
def golden_hashing():
    return load_golden("hashing.json")


def chunks(b, size):
    return [b[i:i + size] for i in range(0, len(b), size)]
