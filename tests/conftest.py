import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


# Slow / external-process cases go last: the driver runs `pytest -x`, so a surprise in one of them cannot hide the results
# of the kernel and codec cases in front of it.
RUN_LAST = ("test_reference_surface_gpu.py", "test_reference_cuda_gpu.py")


def pytest_collection_modifyitems(config, items):
    def rank(it):
        return max((i + 1 for i, tag in enumerate(RUN_LAST) if tag in it.nodeid), default=0)
    items.sort(key=rank)  # stable: the collection order is kept otherwise
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
