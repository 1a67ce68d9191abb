import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """The product path needs the in-tree CUDA library AND a device; fail loudly
    (never skip) when a gpu-marked test runs without them."""
    from sph_mountain_waves_b200 import _capi
    assert _capi.LIB_PATH.exists(), f"{_capi.LIB_PATH} not built"
    assert _have_gpu(), "gpu-marked test selected but no CUDA device is visible"
    return True
