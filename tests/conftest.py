import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed_everything():
    """Every test starts from the same torch seed (CPU and CUDA generators): tests that draw inputs without an explicit
    generator (large-batch tower / FM checks on the device) see the same data on every run, so a tolerance check either
    always passes or always fails."""
    try:
        import torch
        torch.manual_seed(20261017)
    except Exception:
        pass
    yield
