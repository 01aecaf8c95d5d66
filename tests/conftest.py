import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) — run with `-m gpu` on the GPU box")


@pytest.fixture(scope="session")
def libsar():
    """The built C-ABI library; building is part of __graft_entry__.build(), so a missing .so is a hard error."""
    from speech_adapter_routing_b200 import _lib

    if not _lib.LIB_PATH.exists():
        _lib.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda_dev():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (the product path has no CPU fallback)")
    from speech_adapter_routing_b200 import _lib

    assert _lib.lib().sar_device_ok() == 1, "libsar needs an sm_100 device"
    return "cuda:0"
