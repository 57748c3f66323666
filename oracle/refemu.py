"""ctypes binding for oracle/_ref/libbbfft_refemu.so -- TEST INFRASTRUCTURE ONLY.

The library is the UNMODIFIED reference planning + OpenCL-C generator (compiled from the
sources where they lie under /root/reference by oracle/Makefile) driven through a host
work-item emulator (oracle/refemu/refemu.cpp).  It gives "the reference's output on the same
inputs" for any bbfft configuration the reference supports.  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libbbfft_refemu.so")

PVC = "{1024, {16, 32}, 131072, gpu}"  # reference tools/common/info.cpp:9


class Config(C.Structure):
    _fields_ = [
        ("dim", C.c_uint),
        ("shape", C.c_ulong * 5),
        ("fp", C.c_int),
        ("dir", C.c_int),
        ("type", C.c_int),
        ("istride", C.c_ulong * 5),
        ("ostride", C.c_ulong * 5),
        ("cb_source", C.c_char_p),
        ("cb_load", C.c_char_p),
        ("cb_store", C.c_char_p),
    ]


_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libbbfft_refemu.so missing: run `make -C oracle ref`")
        l = C.CDLL(_LIB_PATH)
        l.refemu_last_error.restype = C.c_char_p
        l.refemu_plan_kernel_name.restype = C.c_char_p
        l.refemu_plan_kernel_name.argtypes = [C.c_void_p, C.c_int]
        l.refemu_plan_num_kernels.argtypes = [C.c_void_p]
        l.refemu_plan_create.argtypes = [C.POINTER(Config), C.c_char_p, C.POINTER(C.c_void_p)]
        l.refemu_plan_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.refemu_plan_destroy.argtypes = [C.c_void_p]
        l.refemu_parse_descriptor.argtypes = [C.c_char_p, C.POINTER(Config)]
        l.refemu_to_descriptor.argtypes = [C.POINTER(Config), C.c_char_p, C.c_ulong]
        l.refemu_default_strides.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_ulong),
                                             C.POINTER(C.c_ulong)]
        l.refemu_set_shim_dir(os.path.join(_HERE, "refemu").encode())
        _lib = l
    return _lib


class RefError(RuntimeError):
    pass


class BadConfiguration(RefError):
    pass


def _check(rc):
    if rc != 0:
        msg = lib().refemu_last_error().decode()
        raise (BadConfiguration if rc == 2 else RefError)(msg)


def set_threads(n):
    lib().refemu_set_threads(int(n))


def parse_descriptor(desc):
    c = Config()
    _check(lib().refemu_parse_descriptor(desc.encode(), C.byref(c)))
    return c


def to_descriptor(cfg):
    buf = C.create_string_buffer(512)
    _check(lib().refemu_to_descriptor(C.byref(cfg), buf, 512))
    return buf.value.decode()


def default_strides(cfg, inplace):
    i = (C.c_ulong * 5)()
    o = (C.c_ulong * 5)()
    lib().refemu_default_strides(C.byref(cfg), int(inplace), i, o)
    return list(i), list(o)


def make_config(dim, shape, fp, direction, ttype, istride=None, ostride=None, inplace=True,
                callbacks=None):
    """fp: 4|8, direction: -1|+1, ttype: 0 c2c, 1 r2c, 2 c2r (the reference's enum values)."""
    c = Config()
    c.dim = dim
    for i, s in enumerate(shape):
        c.shape[i] = s
    c.fp, c.dir, c.type = fp, direction, ttype
    di, do = default_strides(c, inplace)
    for i in range(5):
        c.istride[i] = (istride[i] if istride is not None and i < len(istride) else di[i])
        c.ostride[i] = (ostride[i] if ostride is not None and i < len(ostride) else do[i])
    if callbacks is not None:
        src, load, store = callbacks
        c.cb_source = src.encode()
        c.cb_load = load.encode() if load else None
        c.cb_store = store.encode() if store else None
    return c


class Plan:
    def __init__(self, cfg, device_info=PVC):
        self._p = C.c_void_p()
        self._cfg = cfg  # keeps callback strings alive
        _check(lib().refemu_plan_create(C.byref(cfg), device_info.encode(), C.byref(self._p)))

    @property
    def kernel_names(self):
        n = lib().refemu_plan_num_kernels(self._p)
        return [lib().refemu_plan_kernel_name(self._p, i).decode() for i in range(n)]

    def execute(self, inp, out=None):
        """inp/out: C-contiguous numpy buffers; out=None means in-place."""
        assert inp.flags["C_CONTIGUOUS"]
        if out is None:
            out = inp
        assert out.flags["C_CONTIGUOUS"]
        _check(lib().refemu_plan_execute(self._p, inp.ctypes.data, out.ctypes.data))
        return out

    def close(self):
        if self._p:
            lib().refemu_plan_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
