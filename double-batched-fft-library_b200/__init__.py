"""B200-native batched small-FFT path of the Double-Batched FFT Library (bbfft).

Python is only the harness here: device memory, streams and torch.distributed come from
PyTorch; the product is libbbfft_cuda.so (planner + NVRTC JIT + launcher, csrc/) and the
sm_100a kernels in csrc/kernels/bbfft_kernels.cuh, reached through the C ABI of
include/bbfft_cuda.h.  There is no CPU fallback: without the native library import fails.
"""
from .capi import (  # noqa: F401
    BadConfiguration,
    BbfftError,
    Cache,
    Config,
    Plan,
    C2C,
    R2C,
    C2R,
    FORWARD,
    BACKWARD,
    compile_to_cubin,
    default_strides,
    describe,
    describe_chain,
    generate_kernels,
    kernel_header,
    lib,
    make_config,
    parse_descriptor,
    to_descriptor,
)
