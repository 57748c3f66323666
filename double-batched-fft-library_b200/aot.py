"""Ahead-of-time kernel bundles (the role of the reference's tools/aot + cmake/AddAotKernelsToTarget.cmake).

generate_fft_kernels -> CUDA C++ stubs -> nvcc (sm_100a, -lineinfo) -> cubin.  The built-in
bundle (builtin_kernels.cubin next to libbbfft_cuda.so) is registered by the library as an
ahead-of-time cache, so the headline configurations need no NVRTC compile at plan creation.
"""
import os
import subprocess

from . import capi
from .build import ARCH_FLAGS, HERE, KERNELS, NVCC, CXX

def smooth_sizes(lo=2, hi=512, primes=(2, 3, 5, 7)):
    """The reference's benchmark sweep sizes (tools/powers.py -p 4 2 512): 7-smooth numbers."""
    out = []
    for n in range(lo, hi + 1):
        m = n
        for p in primes:
            while m % p == 0:
                m //= p
        if m == 1:
            out.append(n)
    return out


def sweep_k(n, fp, nbytes=1 << 30, m=16):
    """K of the reference benchmark: input tensor of `nbytes` (benchmark/test.hpp:18-30)."""
    return max(1, nbytes // (m * n * 2 * fp))


# FFT descriptors (reference docs/manual/descriptor.rst) of the configurations BASELINE.json names
BUILTIN_DESCRIPTORS = (
    ["scfo64*16384"]
    + ["scfo16.%d*%d" % (n, sweep_k(n, 4)) for n in smooth_sizes()]
    + ["dcfo16.%d*%d" % (n, sweep_k(n, 8)) for n in smooth_sizes()]
    # config 3: r2c / c2r N=256, K=2^20, in- and out-of-place; config 4: 3d fp64 64^3, 2d fp32 128^2
    + ["srfo256*1048576", "srfi256*1048576", "srbo256*1048576", "srbi256*1048576"]
    + ["dcfo64x64x64*64", "scfo128x128*64"]
)


# Sweep sizes whose kernel is built by NVRTC instead of nvcc.  The two front ends hand ptxas different PTX for
# the same stub (e.g. 1400 against 1368 instructions for fp64 N=294) and the result differs by up to 13 % per
# kernel, in either direction, at equal speed over the whole sweep (profiles/r02n_jit_vs_aot.txt: the sweep with
# the nvcc bundle against BBFFT_CUDA_NO_BUILTIN=1, twice each).  Listed here: the sizes where the NVRTC build won
# both repetitions by more than 2 %, and the entries of csrc/wisdom.inc that were chosen from NVRTC-built
# candidates inside the sweep (tools/bench_ab.py).  Arithmetic is unaffected: every multiply is an explicit-
# rounding intrinsic, a GPU test pins nvcc == NVRTC bit for bit.
# fp32 sizes planned with packed adds (X2=1 in csrc/wisdom.inc, profiles/r02y_x2sweep.log) define BBK_F32X2 for their
# translation unit, so they must be alone in it: 196 ... 490 below.
NVRTC_BUILT = {4: (2, 147, 150, 384, 196, 245, 270, 315, 324, 343, 350, 392, 405, 490),
               8: (54, 90, 96, 100, 112, 120, 128, 160, 175, 243, 245, 250, 294, 315, 343, 392, 405, 448, 490)}


def compile_single_nvrtc(descriptor, out_path):
    """One kernel, one translation unit, NVRTC: byte for byte what plan creation would JIT (code generation is
    sensitive to the translation unit as well: the same stub compiled next to others comes out differently)."""
    d = capi.describe(capi.parse_descriptor(descriptor))
    src_copy = os.path.splitext(out_path)[0] + ".cu"
    if not os.path.exists(src_copy) or open(src_copy).read() != d["source"] or not os.path.exists(out_path):
        with open(src_copy, "w") as f:
            f.write(d["source"])
        saved = {k: os.environ.pop(k, None) for k in ("BBFFT_CUDA_KERNEL_CACHE", "BBFFT_CUDA_JIT_LINEINFO")}
        try:
            with open(out_path, "wb") as f:
                f.write(capi.compile_to_cubin(d["source"]))
        finally:
            for k, v in saved.items():
                if v is not None:
                    os.environ[k] = v
    return d["identifier"]


def compile_bundle(descriptors, out_path, verbose=False):
    """descriptors -> cubin at out_path (nvcc, sm_100a); returns the kernel names inside."""
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
        return names, subprocess.Popen(cmd)
    return names, None


def build_builtin(verbose=False, jobs=8):
    """Built-in bundles builtin_kernels_<i>.cubin next to the library (compiled in parallel)."""
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    header = os.path.join(KERNELS, "bbfft_kernels.cuh")
    rtc = ["%scfo16.%d*%d" % ("s" if fp == 4 else "d", n, sweep_k(n, fp)) for fp in (4, 8) for n in NVRTC_BUILT[fp]]
    # config 3: chosen among 817 NVRTC-built candidates per placement (tools/cases_c3.json): ship what was measured
    rtc += ["srfo256*1048576", "srfi256*1048576", "srbo256*1048576", "srbi256*1048576"]
    nvcc_built = [d for d in BUILTIN_DESCRIPTORS if d not in rtc]
    chunks = [nvcc_built[i::jobs] for i in range(jobs)]
    procs = []
    all_names = []
    for i, chunk in enumerate(chunks):
        out = os.path.join(bdir, "builtin_kernels_%d.cubin" % i)
        if os.path.exists(out) and os.path.getmtime(out) < os.path.getmtime(header):
            os.remove(out)
        names, proc = compile_bundle(chunk, out, verbose)
        all_names += names
        procs.append((out, proc))
    for desc in rtc:
        rtc_out = os.path.join(bdir, "builtin_kernels_rtc_%s.cubin" % desc.split("*")[0].replace(".", "_"))
        if os.path.exists(rtc_out) and os.path.getmtime(rtc_out) < os.path.getmtime(header):
            os.remove(rtc_out)
        all_names.append(compile_single_nvrtc(desc, rtc_out))
        procs.append((rtc_out, None))
    for out, proc in procs:
        if proc is not None and proc.wait() != 0:
            raise RuntimeError("nvcc failed for " + out)
    # publish next to the library, dropping stale bundles
    for f in os.listdir(HERE):
        if f.startswith("builtin_kernels") and f.endswith(".cubin"):
            os.remove(os.path.join(HERE, f))
    for out, _ in procs:
        with open(out, "rb") as f, open(os.path.join(HERE, os.path.basename(out)), "wb") as g:
            g.write(f.read())
    return all_names
