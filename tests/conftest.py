import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_pieces():
    """Build the checker-side native code (C oracle, kernel-source CPU harness) if missing.
    libbrdae.so itself is built by __graft_entry__.build() / python -m dae_scipy_b200.build."""
    for d in ("oracle", os.path.join("tests", "host_sim")):
        subprocess.run(["make", "-C", os.path.join(ROOT, d)], check=True, capture_output=True)
    yield
