import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# Ranks that are THREADS of one process (tests/test_gpu_multirank_one_device.py) synchronise through kernels that
# spin on flags written by a peer's kernel.  CUDA's default lazy module loading takes a context-wide lock the first
# time a kernel is launched, which a spinning peer kernel would block forever; eager loading (read once, when the
# process initialises CUDA) removes that first-launch dependency.  Ranks in separate processes are not affected.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
# ... and with 8 such ranks their streams must not be multiplexed onto the default 8 hardware queues: a kernel queued
# behind a peer's spinning barrier kernel would never start
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


# libmvgpu.so binds NCCL at run time by soname (csrc/nccl_dyn.h).  torch brings its own, newer libnccl.so.2: whichever is
# loaded first wins for the whole process, and torch cannot start on the system copy.  Tests that import torch late
# (torch.multiprocessing for the multi-process cases) therefore need torch's copy to be the first one in.
try:
    import torch  # noqa: F401,E402
except Exception:  # pragma: no cover  (CPU-only environments without torch still run the oracle / ABI tests)
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "ref_traces.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def golden_rmat():
    """Reference traces on power-law graphs (tests/golden/make_golden_rmat.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "rmat_traces.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything is built in-tree before any test runs (no-op when up to date): host library, C oracle,
    libmvgpu.so for sm_100a (nvcc cross-compiles without a GPU) and the C++ driver bin/miniVite_b200."""
    import __graft_entry__ as ge
    ge.build()
