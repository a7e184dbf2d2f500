import os
import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def default_code():
    """Test.cpp's code: z=24, N=576, K=432, rate_3_4_b (reference Test.cpp:19-26)."""
    import oracle
    rp, ci, M = oracle.wimax_H(576, "3/4B")
    return dict(M=M, N=576, K=432, row_ptr=rp, col_idx=ci, rate=4)
