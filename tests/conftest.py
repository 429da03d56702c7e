import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_PRESENT = os.path.isdir("/root/reference/code/lib")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device here; GPU parity tests run under gpurun")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.clib.build()
    return oracle


@pytest.fixture
def tuning():
    """tuning(key, value): sets a process-wide tuning switch of libwssdl_b200.so through the C
    ABI (wssdl_set_tuning) for the duration of the test."""
    from wssdl_bus_b200 import _lib
    saved = []

    def set_(key, value):
        saved.append((key, _lib.set_tuning(key, value)))

    yield set_
    for key, prev in reversed(saved):
        _lib.set_tuning(key, prev)
