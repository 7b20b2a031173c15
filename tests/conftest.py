import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# cubins compiled in the build container travel to the GPU box with the working tree
os.environ.setdefault("INFLATOX_CACHE_DIR", os.path.join(ROOT, "tests", ".cubin_cache"))
os.environ.setdefault("INFLATOX_QUIET", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def _has_gpu() -> bool:
    try:
        from inflatox_b200 import _native

        return _native.lib().inflx_device_count() > 0
    except Exception:
        return False


HAS_GPU = _has_gpu()


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
