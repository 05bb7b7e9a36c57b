import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu under gpurun)')


@pytest.fixture(scope='session')
def lib():
    """Builds (if needed) and loads libiiseg.so; CPU-safe (no kernel is launched)."""
    from iterative_inference_segm_b200.csrc.build import build
    from iterative_inference_segm_b200 import _lib
    build()
    return _lib.load()


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('gpu-marked test started without a CUDA device')
    from iterative_inference_segm_b200.csrc.build import build
    build()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device('cuda')
