import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


# Device cases that were written after the round's GPU budget ran out and have only run under the CPU tier's
# kernel emulation so far.  The driver runs `pytest -x`: they go last, so a surprise in one of them cannot hide the
# results of the device-validated cases behind it.  Remove an entry once its first device run is green.  (The capture-
# lane cases are additionally gated: DCVC_B200_TEST_LANES=1, set by tools/round2_first_call.sh — see the test files.)
FIRST_DEVICE_RUN = ("test_hts_gpu.py::test_chunk_roundtrip_state_consistency[2160-3840",
                    "test_ld_gpu.py::test_frame_roundtrip_state_consistency[2160-3840", "test_sequence_gpu.py",
                    "test_hts_gpu.py::test_capture_lanes_bit_identical",
                    "test_codec_gpu.py::test_half_picture_lanes_bit_identical",
                    "test_ld_gpu.py::test_half_picture_lanes_bit_identical",
                    "test_codec_gpu.py::test_decode_one_sync_bit_identical")   # in the order they run


def pytest_collection_modifyitems(config, items):
    def rank(it):
        return max((i + 1 for i, tag in enumerate(FIRST_DEVICE_RUN) if tag in it.nodeid), default=0)
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
