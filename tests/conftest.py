import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_small():
    from tests.util import load_golden
    return load_golden("case_small")


@pytest.fixture(scope="session")
def golden_c1mini():
    from tests.util import load_golden
    return load_golden("case_c1mini")


@pytest.fixture(scope="session")
def golden_c1():
    """BASELINE.json config 1 at its stated size: simulate_tree(10000, 1e-3, 1.4, 0.8) -> 7,108 individuals."""
    from tests.util import load_golden
    return load_golden("case_c1")
