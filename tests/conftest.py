import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def pkg():
    import b200nmpc
    if not b200nmpc._ffi.LIB_PATH.exists():      # a tree that was never built: compile (nvcc cross-compiles without a GPU)
        b200nmpc._ffi.build()
    return b200nmpc
