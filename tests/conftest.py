import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests never silently pass on a CPU box: without a device they are skipped when the
    suite is run unfiltered, and fail if someone selected them with -m gpu."""
    import torch
    if torch.cuda.is_available():
        return
    selected_gpu = "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or "")
    for item in items:
        if "gpu" in item.keywords and not selected_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
