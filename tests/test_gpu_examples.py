"""The reference's example programs re-hosted on the CUDA backend (examples/fft3d,
examples/transpose_fft_transpose): they build in the CPU tier and pass their own checks on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "double-batched-fft-library_b200")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
BIN = os.path.join(ROOT, "examples", "bin")


def _stale(exe, deps):
    return not os.path.exists(exe) or any(os.path.getmtime(exe) < os.path.getmtime(d) for d in deps)


def _build_fft3d():
    src = os.path.join(ROOT, "examples", "fft3d", "fft3d-cuda.cpp")
    exe = os.path.join(BIN, "fft3d-cuda")
    if _stale(exe, [src, os.path.join(LIBDIR, "libbbfft_cuda.so")]):
        os.makedirs(BIN, exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                               "-I" + os.path.join(CUDA_HOME, "include"), src, "-o", exe, "-L" + LIBDIR, "-lbbfft_cuda",
                               "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcufft", "-lcudart", "-Wl,-rpath," + LIBDIR])
    return exe


def _build_tft():
    src = os.path.join(ROOT, "examples", "transpose_fft_transpose", "tft-cuda.cu")
    exe = os.path.join(BIN, "tft-cuda")
    if _stale(exe, [src, os.path.join(LIBDIR, "libbbfft_cuda.so")]):
        os.makedirs(BIN, exist_ok=True)
        subprocess.check_call([os.path.join(CUDA_HOME, "bin", "nvcc"), "-std=c++17", "-O2", "-gencode",
                               "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L" + LIBDIR, "-lbbfft_cuda", "-Xlinker", "-rpath," + LIBDIR])
    return exe


def test_examples_build(pkg):
    assert os.path.exists(_build_fft3d())
    assert os.path.exists(_build_tft())


@pytest.mark.gpu
@pytest.mark.parametrize("flags,dims", [("-r", (64, 64, 64)), ("-ro", (100, 30, 18)), ("-rd", (48, 20, 12)), ("-c", (64, 64, 64)),
                                        ("-cod", (32, 24, 10)), ("-rx", (128, 128, 128)), ("-cdx", (64, 64, 64))])
def test_fft3d_example(pkg, flags, dims):
    """examples/fft3d: the product-of-modes spectrum is checked by the program itself."""
    exe = _build_fft3d()
    r = subprocess.run([exe, flags] + [str(d) for d in dims] + ["2"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-1500:], r.stderr[-1500:])
    assert r.returncode == 0 and "GB/s" in r.stdout


@pytest.mark.gpu
def test_transpose_fft_transpose_example(pkg):
    """examples/transpose_fft_transpose: both routes agree and the double-batched plan is reported."""
    exe = _build_tft()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-1500:])
    assert r.returncode == 0 and r.stdout.count("Speed-up") == 4
