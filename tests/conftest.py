import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "requires_reference: needs /root/reference (authoring container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_harness
    if not ref_harness.available():
        skip = pytest.mark.skip(reason="/root/reference not mounted")
        for it in items:
            if "requires_reference" in it.keywords:
                it.add_marker(skip)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        skip_gpu = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip_gpu)
