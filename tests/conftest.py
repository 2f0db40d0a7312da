import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
def _built_libraries():
    """Make sure the oracle (checker) and the CUDA library exist; both are built in-tree."""
    if not (ROOT / "oracle" / "libmld_oracle.so").exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle")])
    if not (ROOT / "mono_lidar_depth_b200" / "libmld_cuda.so").exists():
        subprocess.check_call(["make", "-C", str(ROOT / "mono_lidar_depth_b200" / "csrc"), "-j8"])
    yield
