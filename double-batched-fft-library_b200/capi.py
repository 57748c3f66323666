"""ctypes binding of include/bbfft_cuda.h (the C ABI of libbbfft_cuda.so).

Mirrors the reference's host interface for the hot path: Config <-> bbfft::configuration
(include/bbfft/configuration.hpp:151-192), Plan <-> make_plan + plan::execute
(include/bbfft/sycl/make_plan.hpp:27-41, include/bbfft/plan.hpp:77-131), Cache <-> jit_cache_all.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbbfft_cuda.so")

C2C, R2C, C2R = 0, 1, 2
FORWARD, BACKWARD = -1, 1
F32, F64 = 4, 8


class BbfftError(RuntimeError):
    pass


class BadConfiguration(BbfftError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("dim", C.c_uint),
        ("shape", C.c_size_t * 5),
        ("fp", C.c_int),
        ("dir", C.c_int),
        ("type", C.c_int),
        ("istride", C.c_size_t * 5),
        ("ostride", C.c_size_t * 5),
        ("cb_source", C.c_char_p),
        ("cb_length", C.c_size_t),
        ("cb_load", C.c_char_p),
        ("cb_store", C.c_char_p),
        ("cb_language", C.c_int),
    ]


class KernelDesc(C.Structure):
    _fields_ = [
        ("identifier", C.c_void_p),
        ("source", C.c_void_p),
        ("twiddle", C.POINTER(C.c_double)),
        ("twiddle_len", C.c_size_t),
        ("grid", C.c_uint64),
        ("threads", C.c_int),
        ("smem_bytes", C.c_size_t),
        ("inplace_unsupported", C.c_int),
        ("fp", C.c_int),
        ("n_stages", C.c_int),
        ("radix", C.c_int * 4),
        ("threads_per_transform", C.c_int),
        ("batch_lanes", C.c_int),
        ("batch_high", C.c_int),
        ("load_staged", C.c_int),
        ("store_staged", C.c_int),
    ]


class ChainDesc(C.Structure):
    _fields_ = [
        ("identifier", C.c_void_p),
        ("source", C.c_void_p),
        ("twiddle", C.POINTER(C.c_double)),
        ("twiddle_len", C.c_size_t),
        ("threads", C.c_int),
        ("smem_bytes", C.c_size_t),
        ("min_blocks", C.c_int),
        ("fp", C.c_int),
        ("n_steps", C.c_int),
        ("step_tile", C.c_int * 3),
        ("per_k", C.c_uint64 * 3),
        ("mult", C.c_uint64 * 3),
        ("step_M", C.c_uint64 * 3),
        ("tw_offset", C.c_int * 3),
        ("uses_tmp", C.c_int),
    ]


_lib = None


def lib():
    """The native library; raises if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise BbfftError(
                "libbbfft_cuda.so is missing: build it with __graft_entry__.build() "
                "(python double-batched-fft-library_b200/build.py)")
        l = C.CDLL(_LIB_PATH)
        l.bbfft_cuda_last_error.restype = C.c_char_p
        l.bbfft_cuda_plan_kernel_name.restype = C.c_char_p
        l.bbfft_cuda_plan_kernel_name.argtypes = [C.c_void_p, C.c_int]
        l.bbfft_cuda_plan_num_kernels.argtypes = [C.c_void_p]
        l.bbfft_cuda_plan_launches.argtypes = [C.c_void_p]
        l.bbfft_cuda_kernel_header.restype = C.c_char_p
        l.bbfft_cuda_default_strides.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(C.c_size_t),
                                                 C.POINTER(C.c_size_t)]
        l.bbfft_cuda_parse_descriptor.argtypes = [C.c_char_p, C.POINTER(Config)]
        l.bbfft_cuda_to_descriptor.argtypes = [C.POINTER(Config), C.c_char_p, C.c_size_t]
        l.bbfft_cuda_cache_create.argtypes = [C.POINTER(C.c_void_p)]
        l.bbfft_cuda_cache_destroy.argtypes = [C.c_void_p]
        l.bbfft_cuda_cache_size.argtypes = [C.c_void_p]
        l.bbfft_cuda_plan_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config), C.c_void_p, C.c_int,
                                             C.c_void_p]
        l.bbfft_cuda_plan_create_tuned.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config), C.c_void_p,
                                                   C.c_int, C.c_void_p, C.c_char_p]
        l.bbfft_cuda_plan_execute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.bbfft_cuda_plan_execute_on.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.bbfft_cuda_plan_execute_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        l.bbfft_cuda_plan_destroy.argtypes = [C.c_void_p]
        l.bbfft_cuda_describe.argtypes = [C.POINTER(Config), C.c_char_p, C.POINTER(KernelDesc)]
        l.bbfft_cuda_desc_free.argtypes = [C.POINTER(KernelDesc)]
        l.bbfft_cuda_desc_free.restype = None
        l.bbfft_cuda_describe_chain.argtypes = [C.POINTER(Config), C.POINTER(ChainDesc)]
        l.bbfft_cuda_chain_desc_free.argtypes = [C.POINTER(ChainDesc)]
        l.bbfft_cuda_chain_desc_free.restype = None
        l.bbfft_cuda_generate_kernels.argtypes = [C.POINTER(Config), C.c_size_t, C.POINTER(C.c_void_p),
                                                  C.POINTER(C.c_void_p)]
        l.bbfft_cuda_compile.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        l.bbfft_cuda_free.argtypes = [C.c_void_p]
        l.bbfft_cuda_free.restype = None
        _lib = l
    return _lib


def _check(rc):
    if rc != 0:
        msg = lib().bbfft_cuda_last_error().decode(errors="replace")
        raise (BadConfiguration if rc == 2 else BbfftError)(msg)


def default_strides(dim, shape, ttype, inplace):
    c = Config()
    c.dim = dim
    for i, s in enumerate(shape):
        c.shape[i] = s
    c.type = ttype
    i_ = (C.c_size_t * 5)()
    o_ = (C.c_size_t * 5)()
    _check(lib().bbfft_cuda_default_strides(C.byref(c), int(inplace), i_, o_))
    return list(i_), list(o_)


def make_config(dim, shape, fp, direction=FORWARD, ttype=C2C, istride=None, ostride=None, inplace=True,
                callbacks=None):
    """shape = [M, N_1..N_dim, K].  Missing strides default like bbfft::configuration's members
    (in-place layout) unless inplace=False.  callbacks = (source, load_name, store_name, language)
    with language "cuda" or "opencl"."""
    c = Config()
    c.dim = dim
    for i, s in enumerate(shape):
        c.shape[i] = s
    c.fp, c.dir, c.type = fp, direction, ttype
    di, do = default_strides(dim, shape, ttype, inplace)
    for i in range(5):
        c.istride[i] = istride[i] if istride is not None and i < len(istride) else di[i]
        c.ostride[i] = ostride[i] if ostride is not None and i < len(ostride) else do[i]
    if callbacks is not None:
        src, load, store = callbacks[:3]
        lang = callbacks[3] if len(callbacks) > 3 else "cuda"
        b = src.encode()
        c._keep = (b, load.encode() if load else None, store.encode() if store else None)
        c.cb_source = c._keep[0]
        c.cb_length = len(b)
        c.cb_load = c._keep[1]
        c.cb_store = c._keep[2]
        c.cb_language = 1 if lang == "cuda" else 0
    return c


def parse_descriptor(desc):
    c = Config()
    _check(lib().bbfft_cuda_parse_descriptor(desc.encode(), C.byref(c)))
    return c


def to_descriptor(cfg):
    buf = C.create_string_buffer(512)
    _check(lib().bbfft_cuda_to_descriptor(C.byref(cfg), buf, 512))
    return buf.value.decode()


def describe(cfg, tune=""):
    """Device-free planning: dict with the kernel identifier, CUDA stub source, launch geometry
    and the twiddle table (float64 pairs) for a 1d configuration."""
    import numpy as np
    d = KernelDesc()
    _check(lib().bbfft_cuda_describe(C.byref(cfg), tune.encode(), C.byref(d)))
    try:
        tw = np.ctypeslib.as_array(d.twiddle, shape=(d.twiddle_len,)).copy()
        out = dict(
            identifier=C.string_at(d.identifier).decode(),
            source=C.string_at(d.source).decode(),
            twiddle=tw,
            grid=int(d.grid),
            threads=int(d.threads),
            smem_bytes=int(d.smem_bytes),
            inplace_unsupported=bool(d.inplace_unsupported),
            fp=int(d.fp),
            radix=list(d.radix)[: min(4, d.n_stages)],
            threads_per_transform=int(d.threads_per_transform),
            batch_lanes=int(d.batch_lanes),
            batch_high=int(d.batch_high),
            load_staged=bool(d.load_staged),
            store_staged=bool(d.store_staged),
        )
    finally:
        lib().bbfft_cuda_desc_free(C.byref(d))
    return out


def describe_chain(cfg):
    """Device-free planning of a 2d/3d configuration as one persistent chain kernel (raises
    BadConfiguration when its steps cannot be chained)."""
    import numpy as np
    d = ChainDesc()
    _check(lib().bbfft_cuda_describe_chain(C.byref(cfg), C.byref(d)))
    try:
        n = int(d.n_steps)
        out = dict(
            identifier=C.string_at(d.identifier).decode(),
            source=C.string_at(d.source).decode(),
            twiddle=np.ctypeslib.as_array(d.twiddle, shape=(max(1, d.twiddle_len),))[: d.twiddle_len].copy(),
            threads=int(d.threads), smem_bytes=int(d.smem_bytes), min_blocks=int(d.min_blocks), fp=int(d.fp),
            n_steps=n, step_tile=list(d.step_tile)[:n], per_k=list(d.per_k)[:n], mult=list(d.mult)[:n],
            step_M=list(d.step_M)[:n], tw_offset=list(d.tw_offset)[:n], uses_tmp=bool(d.uses_tmp),
        )
    finally:
        lib().bbfft_cuda_chain_desc_free(C.byref(d))
    return out


def generate_kernels(cfgs):
    """bbfft::generate_fft_kernels: (source, [kernel names]) for a list of configurations."""
    arr = (Config * len(cfgs))(*cfgs)
    src = C.c_void_p()
    names = C.c_void_p()
    _check(lib().bbfft_cuda_generate_kernels(arr, len(cfgs), C.byref(src), C.byref(names)))
    try:
        return C.string_at(src).decode(), [n for n in C.string_at(names).decode().split("\n") if n]
    finally:
        lib().bbfft_cuda_free(src)
        lib().bbfft_cuda_free(names)


def kernel_header():
    return lib().bbfft_cuda_kernel_header().decode()


def compile_to_cubin(source, arch="sm_100a"):
    binary = C.c_void_p()
    size = C.c_size_t()
    _check(lib().bbfft_cuda_compile(source.encode(), arch.encode(), C.byref(binary), C.byref(size)))
    try:
        return C.string_at(binary, size.value)
    finally:
        lib().bbfft_cuda_free(binary)


class Cache:
    """bbfft::jit_cache_all"""

    def __init__(self):
        self._c = C.c_void_p()
        _check(lib().bbfft_cuda_cache_create(C.byref(self._c)))

    def __len__(self):
        return lib().bbfft_cuda_cache_size(self._c)

    def close(self):
        if self._c:
            lib().bbfft_cuda_cache_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError("expected a torch tensor, numpy array or integer address")


class Plan:
    """make_plan(cfg, stream, cache) + plan::execute.  `stream` is a raw cudaStream_t value
    (e.g. torch.cuda.current_stream().cuda_stream); 0 = default stream."""

    def __init__(self, cfg, stream=0, device=-1, cache=None, tune=""):
        self._p = C.c_void_p()
        self._cfg = cfg
        self._cache = cache
        _check(lib().bbfft_cuda_plan_create_tuned(C.byref(self._p), C.byref(cfg), C.c_void_p(stream), device,
                                                  cache._c if cache is not None else None, tune.encode()))

    @property
    def kernel_names(self):
        n = lib().bbfft_cuda_plan_num_kernels(self._p)
        return [lib().bbfft_cuda_plan_kernel_name(self._p, i).decode() for i in range(n)]

    @property
    def launches_per_execute(self):
        return lib().bbfft_cuda_plan_launches(self._p)

    def execute(self, inp, out=None, stream=None):
        """Asynchronous, stream-ordered.  out=None (or out is inp) -> in-place."""
        pi = _ptr(inp)
        po = pi if out is None else _ptr(out)
        if stream is None:
            _check(lib().bbfft_cuda_plan_execute(self._p, pi, po))
        else:
            _check(lib().bbfft_cuda_plan_execute_on(self._p, pi, po, C.c_void_p(stream)))

    def execute_host(self, inp, out=None):
        """Host buffers (numpy arrays or pinned torch tensors): H2D, transform, D2H, synchronise."""
        if out is None:
            out = inp
        nb = lambda a: a.nbytes if hasattr(a, "nbytes") else a.numel() * a.element_size()
        _check(lib().bbfft_cuda_plan_execute_host(self._p, _ptr(inp), nb(inp), _ptr(out), nb(out)))
        return out

    def close(self):
        if self._p:
            lib().bbfft_cuda_plan_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
