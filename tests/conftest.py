import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def engine():
    """One Engine (C-ABI handle on cuda:0) for the whole GPU session; fails loudly without a GPU."""
    from admm_project_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.close()
