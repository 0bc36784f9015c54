import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _native_library():
    """Make sure cet_pick_b200/libcetpick_sm100a.so is built from the sources in this tree before any test runs
    (a no-op when the in-tree library is up to date).  There is no fallback: if nvcc is missing and so is the
    library, the product raises ImportError and the tests fail loudly."""
    from cet_pick_b200 import build
    try:
        build.build()
    except RuntimeError as e:           # no nvcc on this box: the shipped .so (if any) is used as is
        if not os.path.exists(build.LIB):
            raise
        print(f"conftest: keeping the existing library ({e})")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
