import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available():
    """a usable CUDA device AND the built library (the product has no CPU fallback)"""
    try:
        from lammps_b200 import engine
        return engine.load_library().b200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests need the B200 box: on a machine without a CUDA device (or without the built
    # libb200md.so) they are skipped with the reason instead of failing one by one
    if not any("gpu" in it.keywords for it in items):
        return
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible / libb200md.so not built")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
