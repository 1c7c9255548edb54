import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def native_lib():
    """libb200knn.so, built in-tree if the build box has not done so yet (nvcc cross-compiles without a GPU)."""
    from inclusivegan_b200 import build, load_library
    build.build()
    return load_library()
