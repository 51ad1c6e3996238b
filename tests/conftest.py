import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/surfaces_and_fields")
    skip_ref = pytest.mark.skip(reason="/root/reference not present (GPU box)")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """Path of libcylinder_b200.so; (re)built incrementally when nvcc is present, else the shipped file."""
    from cylinder_b200 import build as b
    try:
        return b.build()
    except RuntimeError:
        if os.path.exists(b.lib_path()):
            return b.lib_path()
        raise


@pytest.fixture(scope="session")
def cuda(built_lib):
    """torch CUDA device for -m gpu tests; fails (not skips) when the extension cannot load."""
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    from cylinder_b200 import _abi
    _abi.lib()
    return torch.device("cuda", 0)
