import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if stale) and load the CUDA library; GPU tests must never run on a fallback."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from clsr_b200 import build, engine
    build.build()
    return engine.load_library()
