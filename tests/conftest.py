import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_device():
    """CUDA ordinal to test on; fails loudly (never skips) when a -m gpu test has no device."""
    import torch

    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    return 0
