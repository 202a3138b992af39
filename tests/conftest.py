import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a CUDA device and the built library: on a box without one they are skipped, not failed (the product
    itself has no CPU path and raises when asked to run there)."""
    try:
        import torch

        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
