import os
import sys

import pytest

# The rank-grid tests run up to 8 ranks INSIDE ONE PROCESS on one GPU (2 streams each).  With the default 8 hardware work
# queues several streams share a queue, and a rank's halo wait kernel (a bounded spin on a peer's flag) at the head of a
# queue can hold back the very kernel that would publish that flag -> a timing-dependent stall that ends in the wait
# kernel's trap.  One process per GPU (the product's launch mode) has 2-3 streams and cannot alias.  Must be set before
# CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


collect_ignore_glob = ["_bin/*", "emul/*", "unit/*"]      # staged reference programs and the emulator source are not test modules


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
