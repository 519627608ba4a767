import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "ref_traces.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything is built in-tree before any test runs (no-op when up to date): host library, C oracle,
    libmvgpu.so for sm_100a (nvcc cross-compiles without a GPU) and the C++ driver bin/miniVite_b200."""
    import __graft_entry__ as ge
    ge.build()
