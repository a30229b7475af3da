import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _gpu_unavailable_reason():
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device (the `gpu` tests need a B200; run them with `pytest -m gpu` under gpurun)"
        if torch.cuda.get_device_capability(0)[0] != 10:
            return "device 0 is not sm_100 (B200)"
        from srgd_b200 import _lib
        if not os.path.exists(_lib.LIB_PATH):
            return f"{_lib.LIB_PATH} has not been built (python -m srgd_b200.build)"
    except Exception as e:                                   # pragma: no cover
        return f"GPU probe failed: {e}"
    return None


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a machine without a B200 (or without the built library) skips the `gpu` tests
    instead of failing them; on a GPU box nothing is skipped, and the product path itself still raises
    loudly when its CUDA library is missing (tests/test_host_logic.py)."""
    if not any("gpu" in item.keywords for item in items):
        return
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
