import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a real B200 (run with -m gpu)")
    config.addinivalue_line("markers", "gpu_pending: GPU test written after the round's GPU budget was spent - never executed on a B200 yet; "
                                       "run with -m gpu_pending, then re-mark as gpu once it has passed there")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords or "gpu_pending" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(1234)
    yield
