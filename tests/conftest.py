import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def built():
    ge.build()
    return True


@pytest.fixture(scope="session")
def pb2(built):
    return ge.load_package()


@pytest.fixture(scope="session")
def scenes(built):
    return ge.load_scenes()


@pytest.fixture(scope="session")
def orc(built):
    return ge.load_oracle()


@pytest.fixture(scope="session")
def gpu(pb2):
    """Initialise device 0; fails loudly (no skip, no fallback) if the extension or the GPU is missing."""
    pb2.init(0)
    return pb2
