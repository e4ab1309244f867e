import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Build (or reuse) the C-ABI library and the oracle's C restatement once per session."""
    from usip_b200 import build as b
    if not os.path.isfile(b.LIB_PATH):
        b.build()
    from oracle import usip_oracle
    usip_oracle.build()
    yield
