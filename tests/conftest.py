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
    """Host library + C oracle are plain gcc builds; make sure they exist before any test runs."""
    import __graft_entry__ as ge
    ge.build_host_only()
