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


KCACHE = os.path.join(ROOT, "kcache")
if os.path.isdir(KCACHE) and os.environ.get("BBFFT_TESTS_NO_KCACHE") != "1":
    # NVRTC results pre-compiled on the CPU-only build box (BBFFT_WARM_CACHE=1 python -m pytest tests -m gpu,
    # see below) travel with the tree: the GPU tier then loads cubins instead of compiling ~2000 kernels.
    # Entries are keyed by a hash of (device header, stub, architecture, options): a stale cache is a miss,
    # never a wrong kernel.
    os.environ.setdefault("BBFFT_CUDA_KERNEL_CACHE", KCACHE)
    os.environ.setdefault("BBFFT_CUDA_JIT_LINEINFO", "0")


def _install_cache_warmer(pkg):
    """BBFFT_WARM_CACHE=1 on a box without a GPU: every 1d / fused-2d plan a GPU test would create is planned
    device-free and its kernel compiled into kcache/, then the test is skipped."""
    import ctypes

    os.makedirs(KCACHE, exist_ok=True)
    os.environ["BBFFT_CUDA_KERNEL_CACHE"] = KCACHE
    os.environ["BBFFT_CUDA_JIT_LINEINFO"] = "0"
    seen = set()

    def warm_one(cfg, tune):
        try:
            d = pkg.describe(cfg, tune or "")
            if d["identifier"] not in seen:
                seen.add(d["identifier"])
                pkg.compile_to_cubin(d["source"])
                # M == 1 real plans may need their unaligned twin (PAIR=0) and spilling kernels their
                # uncapped build: both are compiled on demand on the GPU box, they are rare
        except Exception:
            pass

    def warm(cfg, tune=""):
        if cfg.dim not in (1, 2):
            return
        warm_one(cfg, tune)
        if cfg.dim == 1 and not tune:
            # the test skips at its first plan: also compile the twins its loops would reach (the other
            # placement with default strides, the other direction of a c2c transform)
            shape = [int(cfg.shape[i]) for i in range(3)]
            try:
                for inplace in (False, True):
                    if cfg.type == pkg.C2C:
                        for d in (pkg.FORWARD, pkg.BACKWARD):
                            warm_one(pkg.make_config(1, shape, cfg.fp, d, pkg.C2C, inplace=inplace), "")
                    else:
                        warm_one(pkg.make_config(1, shape, cfg.fp, cfg.dir, cfg.type, inplace=inplace), "")
                        other = pkg.C2R if cfg.type == pkg.R2C else pkg.R2C
                        warm_one(pkg.make_config(1, shape, cfg.fp, -cfg.dir, other, inplace=False), "")
            except Exception:
                pass

    class WarmPlan:
        def __init__(self, cfg, stream=0, device=-1, cache=None, tune=""):
            warm(cfg, tune)
            pytest.skip("cache warm-up run: kernel compiled, nothing executed")

    pkg.Plan = WarmPlan
    try:
        import torch

        class _S:
            cuda_stream = 0

        torch.cuda.current_stream = lambda *a, **k: _S()
    except Exception:
        pass


@pytest.fixture(scope="session")
def pkg():
    """The product package; builds the native library on first use (CPU cross-compile)."""
    lib = os.path.join(ROOT, "double-batched-fft-library_b200", "libbbfft_cuda.so")
    if not os.path.exists(lib):
        importlib.import_module("double-batched-fft-library_b200.build").build_host()
    mod = importlib.import_module("double-batched-fft-library_b200")
    if os.environ.get("BBFFT_WARM_CACHE") == "1":
        _install_cache_warmer(mod)
    return mod


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o
