import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_solver_lib():
    """Builds (if needed) and loads the C-ABI library -- no compute."""
    from dispatches_b200.csrc import build
    build.build()
    from dispatches_b200 import solver
    return solver.load_library()
