"""Ahead-of-time kernel bundles (the role of the reference's tools/aot + cmake/AddAotKernelsToTarget.cmake).

generate_fft_kernels -> CUDA C++ stubs -> nvcc (sm_100a, -lineinfo) -> cubin.  The built-in
bundle (builtin_kernels.cubin next to libbbfft_cuda.so) is registered by the library as an
ahead-of-time cache, so the headline configurations need no NVRTC compile at plan creation.
"""
import os
import subprocess

from . import capi
from .build import ARCH_FLAGS, HERE, KERNELS, NVCC, CXX

# FFT descriptors (reference docs/manual/descriptor.rst) of the configurations BASELINE.json names
BUILTIN_DESCRIPTORS = (
    ["scfo64*16384"]
    + ["scfo16.%d*%d" % (n, (1 << 30) // (16 * n * 8)) for n in (2, 4, 8, 16, 32, 64, 128, 256, 512)]
    + ["dcfo16.%d*%d" % (n, (1 << 30) // (16 * n * 16)) for n in (2, 4, 8, 16, 32, 64, 128, 256, 512)]
)


def compile_bundle(descriptors, out_path, verbose=False):
    """descriptors -> cubin at out_path; returns the kernel names inside."""
    cfgs = [capi.parse_descriptor(d) for d in descriptors]
    source, names = capi.generate_kernels(cfgs)
    cu = os.path.splitext(out_path)[0] + ".cu"
    if not os.path.exists(cu) or open(cu).read() != source or not os.path.exists(out_path):
        with open(cu, "w") as f:
            f.write(source)
        cmd = [NVCC, "-std=c++17", "-O3", "-lineinfo", "-ccbin", CXX] + ARCH_FLAGS + [
            "-I" + KERNELS, "-cubin", "-o", out_path, cu]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return names


def build_builtin(verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    out = os.path.join(HERE, "builtin_kernels.cubin")
    names = compile_bundle(BUILTIN_DESCRIPTORS, os.path.join(HERE, "build", "builtin_kernels.cubin"), verbose)
    # publish atomically next to the library
    tmp = out + ".tmp"
    with open(os.path.join(HERE, "build", "builtin_kernels.cubin"), "rb") as f, open(tmp, "wb") as g:
        g.write(f.read())
    os.replace(tmp, out)
    return names
