"""tests/emu/emu.py -- TEST INFRASTRUCTURE ONLY.

Runs a planned bbfft CUDA kernel on the CPU: the stub the planner generated and the device
header are compiled with g++ (-DBBFFT_EMU) against cuda_emu.hpp and executed one fiber per
CUDA thread by emu_runner.cpp.  This lets the CPU-only test tier check every index map of the
real kernel source against the oracle.  Never used by the product path.
"""
import ctypes as C
import hashlib
import importlib
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_KERNELS = os.path.join(_ROOT, "double-batched-fft-library_b200", "csrc", "kernels")
_CACHE = os.path.join(tempfile.gettempdir(), "bbfft_emu_cache_%d" % os.getuid())

pkg = importlib.import_module("double-batched-fft-library_b200")


class Args(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("out", C.c_void_p), ("tw", C.c_void_p), ("K", C.c_ulonglong),
                ("M", C.c_ulonglong), ("is1", C.c_longlong), ("is2", C.c_longlong), ("os1", C.c_longlong),
                ("os2", C.c_longlong), ("pf", C.c_ulonglong)]


RACECHECK = os.environ.get("BBFFT_EMU_RACECHECK", "0") == "1"


def _compile(desc):
    os.makedirs(_CACHE, exist_ok=True)
    # BBFFT_EMU_KERNELS_DIR: a directory with another bbfft_kernels.cuh (the mutation tests of the emulator itself)
    kernels = os.environ.get("BBFFT_EMU_KERNELS_DIR") or _KERNELS
    hdr = open(os.path.join(kernels, "bbfft_kernels.cuh")).read()
    runner = open(os.path.join(_HERE, "emu_runner.cpp")).read()
    shim = open(os.path.join(_HERE, "cuda_emu.hpp")).read()
    key = hashlib.sha1((desc["source"] + hdr + runner + shim + str(RACECHECK)).encode()).hexdigest()[:20]
    so = os.path.join(_CACHE, key + ".so")
    if not os.path.exists(so):
        stub = os.path.join(_CACHE, key + ".%d.cpp" % os.getpid())
        with open(stub, "w") as f:
            f.write(desc["source"])
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        extra = ["-DBBFFT_EMU_CHAIN"] if desc["source"].startswith("#define BBFFT_EMU_CHAIN") else []
        if RACECHECK:
            extra.append("-DBBFFT_EMU_RACECHECK")
        cmd = [cxx, "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-DBBFFT_EMU"] + extra + [
               "-DBBFFT_EMU_KERNEL=" + desc["identifier"], "-I" + _HERE, "-I" + kernels, "-include",
               os.path.join(_HERE, "cuda_emu.hpp"), stub, os.path.join(_HERE, "emu_runner.cpp"), "-o",
               so + ".%d.tmp" % os.getpid()]  # (pytest-xdist workers may build the same kernel at once)
        subprocess.check_call(cmd)
        os.replace(so + ".%d.tmp" % os.getpid(), so)
    return C.CDLL(so)


def run(cfg, inp, out=None, tune="", grid=None):
    """Execute the planned kernel for a 1d (or fused 2d) `cfg` on numpy buffers (out=None -> in place).
    `grid`: number of CTAs to launch instead of the planned grid (persistent tile kernels walk the tiles)."""
    if cfg.dim == 2 and "CL=" not in tune:
        # thread-block clusters (tiles split over several CTAs, distributed shared memory) are exercised on
        # the GPU tier only: the emulator runs one CTA at a time
        tune = (tune + "," if tune else "") + "CL=1"
    desc = pkg.describe(cfg, tune)
    lib = _compile(desc)
    lib.emu_launch.argtypes = [C.POINTER(Args), C.c_ulonglong, C.c_int, C.c_ulong]
    tw = desc["twiddle"].astype(np.float32 if desc["fp"] == 4 else np.float64)
    if out is None:
        out = inp
    # 1d: K slices; fused 2d: the kernel's K argument counts tiles (= CTAs)
    nk = cfg.shape[2] if cfg.dim == 1 else desc["grid"]
    a = Args(inp.ctypes.data, out.ctypes.data, tw.ctypes.data, nk, cfg.shape[0], cfg.istride[1],
             cfg.istride[2], cfg.ostride[1], cfg.ostride[2])
    rc = lib.emu_launch(C.byref(a), grid or desc["grid"], desc["threads"], desc["smem_bytes"])
    if rc != 0:
        raise RuntimeError("emulated kernel failed (rc=%d): %s" % (rc, {2: "divergent barriers", 4: "shared-memory race (see stderr)"}.get(rc, "device-side failure")))
    return out, desc


class ChainArgs(C.Structure):
    _fields_ = [("step", Args * 3), ("done", C.c_void_p), ("epoch", C.c_ulonglong), ("K", C.c_ulonglong),
                ("kblock", C.c_ulonglong)]


def run_chain(cfg, inp, out, kblock=2, epochs=1):
    """Execute the persistent chain kernel of a 2d/3d `cfg` with ONE emulated CTA (it walks all work
    items in dependency order, so a wait that would spin on the GPU is a test failure)."""
    desc = pkg.describe_chain(cfg)
    desc["source"] = "#define BBFFT_EMU_CHAIN 1\n" + desc["source"]
    lib = _compile(desc)
    lib.emu_launch.argtypes = [C.POINTER(ChainArgs), C.c_ulonglong, C.c_int, C.c_ulong]
    rdt = np.float32 if desc["fp"] == 4 else np.float64
    tw = desc["twiddle"].astype(rdt)
    K = cfg.shape[cfg.dim + 1]
    n = desc["n_steps"]
    done = np.zeros(n * K, dtype=np.uint64)
    assert not desc["uses_tmp"], "emulated chains route intermediates through `out`"
    for epoch in range(1, epochs + 1):
        a = ChainArgs()
        for d in range(n):
            src = inp if d == 0 else out
            a.step[d] = Args(src.ctypes.data, out.ctypes.data, tw.ctypes.data + desc["tw_offset"][d] * 2 * tw.itemsize,
                             desc["mult"][d] * K, desc["step_M"][d], 0, 0, 0, 0)
        a.done, a.epoch, a.K, a.kblock = done.ctypes.data, epoch, K, kblock
        rc = lib.emu_launch(C.byref(a), 1, desc["threads"], desc["smem_bytes"])
        if rc != 0:
            raise RuntimeError("emulated chain kernel failed (rc=%d)" % rc)
    return out, desc, done
