import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_everything():
    """Build the oracle (C restatement) and, where it is stale, the CUDA extension (nvcc cross-compiles on CPU)."""
    from oracle import build as ob
    ob.build_oracle()
    import shutil
    if shutil.which("nvcc"):
        from fplll_b200 import build as pb
        pb.build_all()
    yield
