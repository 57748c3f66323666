import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package; builds the native library on first use (CPU cross-compile)."""
    lib = os.path.join(ROOT, "double-batched-fft-library_b200", "libbbfft_cuda.so")
    if not os.path.exists(lib):
        importlib.import_module("double-batched-fft-library_b200.build").build_host()
    return importlib.import_module("double-batched-fft-library_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o
