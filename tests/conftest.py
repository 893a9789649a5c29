import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "gpu_next: opt-in code paths written without GPU time left in the round; not part of "
                            "-m gpu until they have run green on a B200 once (python -m pytest tests -m gpu_next)")
    try:  # the torch ops used as fp32 references in the kernel tests must not run in TF32
        import torch

        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:
        pass


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords or "gpu_next" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib_built():
    import __graft_entry__

    __graft_entry__.build()
    return True
