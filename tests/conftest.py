import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must not silently pass on a box without a GPU: they are only ever selected with
    `-m gpu`; if someone runs them without a device they fail inside the product (no CPU fallback)."""
    return


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def built_library():
    """The in-tree CUDA library; built on demand (nvcc cross-compiles sm_100a without a GPU)."""
    from mansy_immersivevideostreaming_b200.build import build_library
    return build_library()
