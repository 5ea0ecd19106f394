import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the compiled reference in oracle/_ref")


def _cuda_usable():
    """A CUDA device and the built library: what every `gpu` test needs."""
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(repo, "cammiq_b200", "libcammiq_gpu.so")):
        return False, "cammiq_b200/libcammiq_gpu.so is not built"
    try:
        import ctypes
        rt = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if rt.cuInit(0) != 0 or rt.cuDeviceGetCount(ctypes.byref(n)) != 0 or n.value < 1:
            return False, "no CUDA device"
    except OSError:
        return False, "no CUDA driver"
    return True, ""


def pytest_collection_modifyitems(config, items):
    """`pytest tests/` on a box without a GPU (or without oracle/_ref) skips what cannot run
    instead of erroring; `-m gpu` on the GPU box is unaffected."""
    ok, why = _cuda_usable()
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    have_ref = os.access(os.path.join(repo, "oracle", "_ref", "ref_harness"), os.X_OK)
    for item in items:
        if not ok and "gpu" in item.keywords:
            item.add_marker(pytest.mark.skip(reason=why))
        if not have_ref and "ref" in item.keywords:
            item.add_marker(pytest.mark.skip(reason="oracle/_ref not built"))
