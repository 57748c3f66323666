"""ctypes binding for oracle/libbbfft_oracle.so -- TEST INFRASTRUCTURE ONLY.

oracle_dft    long-double direct DFT over the bbfft tensor layout (ground truth)
oracle_bbfft  reference-structured restatement in working precision
plus the integer helpers the reference's host tests pin (factor, scrambler).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbbfft_oracle.so")

C2C, R2C, C2R = 0, 1, 2
FORWARD, BACKWARD = -1, 1


class Config(C.Structure):
    _fields_ = [
        ("dim", C.c_uint),
        ("shape", C.c_size_t * 5),
        ("fp", C.c_int),
        ("dir", C.c_int),
        ("type", C.c_int),
        ("istride", C.c_size_t * 5),
        ("ostride", C.c_size_t * 5),
    ]


class Device(C.Structure):
    _fields_ = [
        ("max_work_group_size", C.c_size_t),
        ("min_subgroup_size", C.c_size_t),
        ("max_subgroup_size", C.c_size_t),
        ("local_memory_size", C.c_size_t),
        ("is_cpu", C.c_int),
    ]


PVC = Device(1024, 16, 32, 131072, 0)  # reference tools/common/info.cpp:9

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        l = C.CDLL(_LIB_PATH)
        l.oracle_dft.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p]
        l.oracle_bbfft.argtypes = [C.POINTER(Config), C.POINTER(Device), C.c_void_p, C.c_void_p]
        l.oracle_default_strides.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_size_t),
                                             C.POINTER(C.c_size_t)]
        l.oracle_trial_division.argtypes = [C.c_int, C.POINTER(C.c_int)]
        l.oracle_factor.argtypes = [C.c_uint, C.c_uint, C.POINTER(C.c_uint)]
        l.oracle_scramble.restype = C.c_long
        l.oracle_scramble.argtypes = [C.c_long, C.POINTER(C.c_int), C.c_int]
        l.oracle_unscramble.restype = C.c_long
        l.oracle_unscramble.argtypes = [C.c_long, C.POINTER(C.c_int), C.c_int]
        l.oracle_select_1d.argtypes = [C.POINTER(Config), C.POINTER(Device), C.POINTER(C.c_uint),
                                       C.POINTER(C.c_int)]
        _lib = l
    return _lib


def default_strides(dim, shape, ttype, inplace):
    c = Config()
    c.dim = dim
    for i, s in enumerate(shape):
        c.shape[i] = s
    c.type = ttype
    i_ = (C.c_size_t * 5)()
    o_ = (C.c_size_t * 5)()
    lib().oracle_default_strides(C.byref(c), int(inplace), i_, o_)
    return list(i_), list(o_)


def make_config(dim, shape, fp, direction, ttype, istride=None, ostride=None, inplace=True):
    c = Config()
    c.dim = dim
    for i, s in enumerate(shape):
        c.shape[i] = s
    c.fp, c.dir, c.type = fp, direction, ttype
    di, do = default_strides(dim, shape, ttype, inplace)
    for i in range(5):
        c.istride[i] = istride[i] if istride is not None and i < len(istride) else di[i]
        c.ostride[i] = ostride[i] if ostride is not None and i < len(ostride) else do[i]
    return c


def _run(fn, cfg, inp, out, *extra):
    assert inp.flags["C_CONTIGUOUS"]
    if out is None:
        out = inp
    assert out.flags["C_CONTIGUOUS"]
    rc = fn(C.byref(cfg), *extra, inp.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle: unsupported configuration (rc=%d)" % rc)
    return out


def dft(cfg, inp, out=None):
    """Ground truth: long double direct DFT (out=None -> in-place)."""
    return _run(lib().oracle_dft, cfg, inp, out)


def bbfft(cfg, inp, out=None, device=PVC):
    """Reference-structured restatement (same algorithm selection as the reference on `device`)."""
    return _run(lib().oracle_bbfft, cfg, inp, out, C.byref(device))


def trial_division(n):
    buf = (C.c_int * 64)()
    cnt = lib().oracle_trial_division(n, buf)
    return list(buf[:cnt])


def factor(n, index):
    buf = (C.c_uint * max(index, 1))()
    lib().oracle_factor(n, index, buf)
    return list(buf[:index])


def scramble(i, factors):
    f = (C.c_int * len(factors))(*factors)
    return lib().oracle_scramble(i, f, len(factors))


def unscramble(i, factors):
    f = (C.c_int * len(factors))(*factors)
    return lib().oracle_unscramble(i, f, len(factors))


def select_1d(cfg, device=PVC):
    f = (C.c_uint * 8)()
    n = C.c_int()
    path = lib().oracle_select_1d(C.byref(cfg), C.byref(device), f, C.byref(n))
    return ("f2fft" if path else "sbfft"), list(f[: n.value])


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) else np.float64).ravel()
    b = np.asarray(b).astype(np.complex128 if np.iscomplexobj(b) else np.float64).ravel()
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (nb if nb > 0 else 1.0))
