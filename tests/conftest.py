import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref_ext():
    """The UNMODIFIED reference CUDA rasterizer built into oracle/_ref (oracle/build_ref.sh)."""
    from tests import refimpl
    mod = refimpl.load_reference()
    if mod is None:
        pytest.skip("oracle/_ref is not built (run oracle/build_ref.sh where /root/reference exists)")
    return mod
