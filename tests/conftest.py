import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _have_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree native pieces exist (cheap when up to date)."""
    from legitengine_b200 import _build

    _build.build_scene()
    if not (ROOT / "oracle" / "liboracle_port.so").exists():
        _build.build_oracle()
    if not (ROOT / "legitengine_b200" / "lib" / "liblgcu.so").exists():
        import shutil

        if shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists():
            _build.build_cuda()  # cross-compiles without a GPU; the ABI tests then load it
        # without nvcc the CPU tests that need the library skip themselves (abi.LibraryMissing), the others run
    yield
