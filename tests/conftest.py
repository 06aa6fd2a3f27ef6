import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    import opf_graph_neural_solver_b200 as pkg
    if not os.path.exists(pkg.library_path()):
        pkg.build_library()
    return pkg.load_library()
