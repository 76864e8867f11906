import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with `-m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` selects GPU tests explicitly; without a usable device they are skipped rather than erroring
    try:
        import torch
        has_gpu = torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no sm_100 GPU visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
