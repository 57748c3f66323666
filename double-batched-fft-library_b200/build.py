"""Build the native parts of the B200 bbfft backend in-tree.

  libbbfft_cuda.so        host library (planner, NVRTC JIT, CUDA runtime launcher, C ABI, C++ API)
  builtin_kernels.cubin   nvcc-compiled (sm_100a) bundle of the headline kernels, registered as
                          a built-in ahead-of-time cache at library load

Run as a script or call build().  No GPU is needed (nvcc / NVRTC cross-compile).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
KERNELS = os.path.join(CSRC, "kernels")
LIB = os.path.join(HERE, "libbbfft_cuda.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
# the image's default g++ wrapper links libstdc++ statically, which clashes inside dlopen'ed
# libraries; use the system compiler
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]

HOST_SOURCES = ["core.cpp", "planner.cpp", "runtime.cpp", "plan.cpp", "c_abi.cpp", "kernel_header.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def embed_header():
    src = os.path.join(KERNELS, "bbfft_kernels.cuh")
    dst = os.path.join(CSRC, "kernel_header_embed.inc")
    text = open(src).read()
    # split into raw-string chunks (compilers limit the length of one literal)
    chunks = []
    step = 8000
    for i in range(0, len(text), step):
        chunks.append('R"BBFFTHDR(' + text[i:i + step] + ')BBFFTHDR"')
    body = "\n".join(chunks) + "\n"
    if not os.path.exists(dst) or open(dst).read() != body:
        with open(dst, "w") as f:
            f.write(body)
    return dst


def build_host(verbose=False):
    inc = embed_header()
    srcs = [os.path.join(CSRC, s) for s in HOST_SOURCES]
    hdrs = [inc] + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".hpp", ".inc"))]
    hdrs += [os.path.join(ROOT, "include", "bbfft_cuda.h"), os.path.join(ROOT, "include", "bbfft", "api.hpp")]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        if _newer(o, [s] + hdrs):
            cmd = [CXX, "-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wextra",
                   "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I" + os.path.join(CUDA_HOME, "include"),
                   "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("compile failed: " + " ".join(cmd))
    if _newer(LIB, objs):
        cmd = [CXX, "-shared", "-o", LIB] + objs + [
            "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcudart_static", "-ldl", "-lrt", "-lpthread",
            "-Wl,--exclude-libs,ALL"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


def build_tools(verbose=False):
    """bbfft-aot-generate / bbfft-offline-generate / bbfft-device-info (tools/native/bbfft_tools.cpp)."""
    src = os.path.join(ROOT, "tools", "native", "bbfft_tools.cpp")
    bindir = os.path.join(ROOT, "tools", "bin")
    os.makedirs(bindir, exist_ok=True)
    main = os.path.join(bindir, "bbfft-aot-generate")
    if _newer(main, [src, LIB]):
        cmd = [CXX, "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA_HOME, "include"),
               src, "-o", main, "-L" + HERE, "-lbbfft_cuda", "-Wl,-rpath," + HERE]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    # native benchmark driver with cuFFT beside it (benchmark/main-cuda.cu)
    bench_src = os.path.join(ROOT, "benchmark", "main-cuda.cu")
    bench = os.path.join(bindir, "bbfft-bench")
    if os.path.exists(bench_src) and _newer(bench, [bench_src, LIB]):
        cmd = [NVCC, "-std=c++17", "-O2", "-lineinfo", "-ccbin", CXX] + ARCH_FLAGS + [
            "-I" + os.path.join(ROOT, "include"), bench_src, "-o", bench, "-L" + HERE, "-lbbfft_cuda", "-lcufft",
            "-Xlinker", "-rpath," + HERE]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    for alias in ("bbfft-offline-generate", "bbfft-device-info"):
        dst = os.path.join(bindir, alias)
        if os.path.lexists(dst):
            os.remove(dst)
        os.symlink("bbfft-aot-generate", dst)
    return bindir


def build(verbose=False):
    build_host(verbose)
    build_tools(verbose)
    from . import aot  # noqa: deferred, needs the library
    aot.build_builtin(verbose)
    return LIB


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    build_host(verbose=True)
    build_tools(verbose=True)
