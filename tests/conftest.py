import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from hoigen_b200 import _cabi
    _cabi.init(0)
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def enc_state():
    from hoigen_b200 import synthetic as S
    return S.make_encoder_state(0)
