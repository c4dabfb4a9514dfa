import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return load


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device AND the built library: skip them (instead of failing) where either is missing."""
    import torch
    from excel_b200 import _lib
    if torch.cuda.is_available() and os.path.exists(_lib.LIB_PATH):
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and excel_b200/lib/libexcel_b200.so (run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
