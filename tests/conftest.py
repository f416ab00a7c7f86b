import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    n = None
    for item in items:
        if "gpu" in item.keywords or "multigpu" in item.keywords:
            if n is None:
                n = _ngpu()
            if n == 0:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))
            elif "multigpu" in item.keywords and n < 2:
                item.add_marker(pytest.mark.skip(reason="needs >= 2 CUDA devices"))
