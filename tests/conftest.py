"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` = oracle vs golden vectors, host logic, C-ABI load/exports (runs in the build
container, no GPU); `-m gpu` = the parity tests proper, through the C-ABI, on a B200.
"""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _checkers import ensure_checkers_fresh
    ensure_checkers_fresh()


def _have_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        if cuda.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        return cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
