import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_GPU_STATE = {}


def _gpu_unavailable_reason():
    """None when libfrx_b200.so loads and a context can be created on device 0, else why not (cached)."""
    if "reason" not in _GPU_STATE:
        try:
            from frenetix_motion_planner_b200 import _capi
            _capi.Handler(0).close()
            _GPU_STATE["reason"] = None
        except Exception as e:      # missing library or no CUDA device
            _GPU_STATE["reason"] = f"no usable CUDA device / library: {e}"
    return _GPU_STATE["reason"]


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU-only box skips the gpu-marked tests instead of failing in them.  An explicit `-m gpu` run
    (the GPU box) is NOT softened: there a missing device or library must fail loudly."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu") is not None]
    if not gpu_items:
        return
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for it in gpu_items:
        it.add_marker(skip)
