"""CPU tier: the host side of the product (no GPU): the C-ABI library loads and exports every
symbol include/bbfft_cuda.h declares, descriptors / strides / planner behave like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "bbfft_cuda.h")).read()
    names = set(re.findall(r"\b(bbfft_cuda_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 18
    lib = ctypes.CDLL(os.path.join(ROOT, "double-batched-fft-library_b200", "libbbfft_cuda.so"))
    for n in sorted(names):
        assert hasattr(lib, n), "missing export " + n


def test_default_strides(pkg):
    # reference src/base/configuration.cpp:24-56, docs/manual/data-layout.rst
    assert pkg.default_strides(1, [4, 10, 3], pkg.C2C, True) == ([1, 4, 40, 0, 0], [1, 4, 40, 0, 0])
    assert pkg.default_strides(1, [4, 10, 3], pkg.R2C, True) == ([1, 4, 48, 0, 0], [1, 4, 24, 0, 0])
    assert pkg.default_strides(1, [4, 10, 3], pkg.R2C, False) == ([1, 4, 40, 0, 0], [1, 4, 24, 0, 0])
    assert pkg.default_strides(1, [4, 10, 3], pkg.C2R, False) == ([1, 4, 24, 0, 0], [1, 4, 40, 0, 0])
    assert pkg.default_strides(3, [2, 5, 6, 7, 3], pkg.R2C, True) == ([1, 2, 12, 72, 504], [1, 2, 6, 36, 252])
    assert pkg.default_strides(1, [1, 256, 8], pkg.R2C, True)[0][:3] == [1, 1, 258]


@pytest.mark.parametrize("desc,dim,shape,fp,d,t", [
    ("srfi5", 1, [1, 5, 1], 4, -1, 1),
    ("dcbi4*5", 1, [1, 4, 5], 8, 1, 0),
    ("dcbi4.5", 1, [4, 5, 1], 8, 1, 0),
    ("drbo4.5*6", 1, [4, 5, 6], 8, 1, 2),
    ("drfo5x6x7", 3, [1, 5, 6, 7, 1], 8, -1, 1),
    ("srbo4.5x6*7", 2, [4, 5, 6, 7], 4, 1, 2),
    ("scfo16*32i1,1,20", 1, [1, 16, 32], 4, -1, 0),
    ("scfo16*32i1,1,20o1,2,32", 1, [1, 16, 32], 4, -1, 0),
    ("scfi16*32i1,1,20o1,1,20", 1, [1, 16, 32], 4, -1, 0),
])
def test_descriptor_round_trip(pkg, desc, dim, shape, fp, d, t):
    # reference test/parser.cpp:33-126
    c = pkg.parse_descriptor(desc)
    assert c.dim == dim and list(c.shape)[: len(shape)] == shape
    assert (c.fp, c.dir, c.type) == (fp, d, t)
    assert pkg.to_descriptor(c) == desc


@pytest.mark.parametrize("bad", ["", "x", "scf", "scfo", "scfox", "scfo5x", "scfo1.2.3", "scfo2x3x4x5", "scfo5i1,2", "scfo5q"])
def test_descriptor_malformed(pkg, bad):
    with pytest.raises(pkg.BbfftError):
        pkg.parse_descriptor(bad)


def test_identifier_is_independent_of_k(pkg):
    # K is a run-time argument, not part of the cache key (reference examples/cache/main.cpp:50-52)
    a = pkg.describe(pkg.make_config(1, [16, 64, 1000], 4))
    b = pkg.describe(pkg.make_config(1, [16, 64, 5000], 4))
    assert a["identifier"] == b["identifier"]
    c = pkg.describe(pkg.make_config(1, [16, 64, 1000], 8))
    assert c["identifier"] != a["identifier"]


def test_planner_covers_the_sweep(pkg):
    aot = __import__("importlib").import_module("double-batched-fft-library_b200.aot")
    sizes = aot.smooth_sizes()
    assert len(sizes) == 105 and sizes[:6] == [2, 3, 4, 5, 6, 7] and sizes[-1] == 512
    for fp in (4, 8):
        for n in sizes:
            d = pkg.describe(pkg.make_config(1, [16, n, 64], fp, inplace=False))
            prod = int(np.prod(d["radix"]))
            assert prod == n
            assert d["threads"] <= 1024 and d["smem_bytes"] <= 227 * 1024
            assert d["grid"] >= 1


def test_bad_configurations(pkg):
    # stride[0] != 1 -> bad_configuration; r2c must be forward (reference src/base/configuration.cpp:107-111)
    c = pkg.make_config(1, [4, 8, 2], 4, istride=[2, 8, 64], ostride=[1, 4, 32])
    with pytest.raises(pkg.BadConfiguration):
        pkg.describe(c)
    c = pkg.make_config(1, [4, 8, 2], 4, pkg.BACKWARD, pkg.R2C)
    with pytest.raises(pkg.BadConfiguration):
        pkg.describe(c)


def test_generate_kernels_and_nvrtc_cross_compile(pkg):
    """generate_fft_kernels emits one stub per distinct kernel; NVRTC builds it for sm_100a
    without a GPU (the product's JIT path, minus the module load)."""
    cfgs = [pkg.parse_descriptor(d) for d in ("scfo16.64*100", "scfo16.64*200", "dcfo8x8*4")]
    src, names = pkg.generate_kernels(cfgs)
    assert len(names) == len(set(names)) == 3  # same 1d kernel for both K; two passes for the 2d plan
    for n in names:
        assert ("extern \"C\" BBK_GLOBAL" in src) and n in src
    cubin = pkg.compile_to_cubin(src)
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 4096
    for n in names:
        assert n.encode() in cubin


def test_register_cap_models_the_four_register_file_partitions(pkg):
    """reg_cap (planner.cpp): the __maxnreg__ the planner emits must make `mb` CTAs resident under
    the per-partition register model -- 7-warp CTAs at 136 registers only fit once per SM although
    2 x 224 x 136 < 65536 (measured: the 2x slowdown of profiles/r01b_*)."""
    d = pkg.describe(pkg.make_config(1, [16, 135, 64], 8, inplace=False), "R=9x15,T=9,ML=8,BH=3,MB=2")
    m = re.search(r"BBK_KERNEL\(\d+, (\d+)\)", d["source"])
    assert m and int(m.group(1)) == 128 and d["threads"] == 216
    d = pkg.describe(pkg.make_config(1, [16, 135, 64], 8, inplace=False), "R=9x15,T=9,ML=8,BH=3,MB=1")
    assert int(re.search(r"BBK_KERNEL\(\d+, (\d+)\)", d["source"]).group(1)) == 255


def test_builtin_bundle_meets_planned_occupancy(pkg):
    """Every kernel of the nvcc-built ahead-of-time bundle fits the occupancy the planner asked
    for (identifier field _mb<k>) under the per-partition register model, and spills at most a few
    words.  Runs without a GPU (cuobjdump reads the cubins)."""
    import glob
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    cubins = sorted(glob.glob(os.path.join(ROOT, "double-batched-fft-library_b200", "builtin_kernels_*.cubin")))
    assert cubins, "build() has not produced the built-in bundle"
    seen = 0
    for f in cubins:
        out = subprocess.run(["cuobjdump", "-res-usage", f], capture_output=True, text=True).stdout
        for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
            name, reg, stack = m.group(1), int(m.group(2)), int(m.group(3))
            seen += 1
            mb = int(re.search(r"_mb(\d+)_", name).group(1))
            th = re.search(r"_th(\d+)_", name)
            if th:
                threads = int(th.group(1))
            else:
                t, ml, bh = (int(re.search(r"_%s(\d+)_" % k, name).group(1)) for k in ("T", "ML", "BH"))
                threads = t * ml * bh
            warps = (threads + 31) // 32
            warps_per_partition = 16384 // (((reg + 7) // 8 * 8) * 32)
            assert warps_per_partition // ((warps + 3) // 4) >= mb, (name, reg)
            assert stack <= 128, (name, stack)
    assert seen >= 216


def test_persistent_kernel_cache(pkg, tmp_path, monkeypatch):
    """BBFFT_CUDA_KERNEL_CACHE: an NVRTC result is stored under a hash of (header, stub, arch, options)
    and served from disk afterwards; a different stub or option set gets a different file."""
    import os
    d = pkg.describe(pkg.parse_descriptor("scfo16.12*7"))
    monkeypatch.delenv("BBFFT_CUDA_KERNEL_CACHE", raising=False)  # (conftest points the suite at kcache/ when it exists)
    monkeypatch.setenv("BBFFT_CUDA_JIT_LINEINFO", "1")
    plain = pkg.compile_to_cubin(d["source"])
    monkeypatch.setenv("BBFFT_CUDA_KERNEL_CACHE", str(tmp_path))
    first = pkg.compile_to_cubin(d["source"])
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].endswith(".cubin") and first == plain
    # a hit returns the stored bytes: make them recognisable
    path = os.path.join(tmp_path, files[0])
    with open(path, "rb") as f:
        stored = f.read()
    assert stored == first
    with open(path, "wb") as f:
        f.write(stored + b"\0")
    assert pkg.compile_to_cubin(d["source"]) == stored + b"\0"
    # other stub, other options: other entries
    pkg.compile_to_cubin(pkg.describe(pkg.parse_descriptor("scfo16.12*7"), "R=3x4,T=3")["source"])
    monkeypatch.setenv("BBFFT_CUDA_JIT_LINEINFO", "0")
    small = pkg.compile_to_cubin(d["source"])
    assert len(os.listdir(tmp_path)) == 3 and len(small) < len(first)


def test_every_length_plans(pkg):
    """Like the reference, no N is refused: primes beyond the in-register butterflies become
    direct-DFT stages (radix > 31 in the identifier) and always sit in a stage of their own."""
    for n in (37, 53, 67, 101, 127, 254, 379, 424, 509, 2 * 251):
        for fp in (4, 8):
            d = pkg.describe(pkg.make_config(1, [16, n, 8], fp, pkg.FORWARD, pkg.C2C, inplace=False))
            prod = 1
            for r in d["radix"]:
                prod *= r
                big = [p for p in range(32, r + 1) if r % p == 0 and all(p % q for q in range(2, int(p ** 0.5) + 1))]
                assert not big or big == [r], (n, d["radix"])
            assert prod == n and d["smem_bytes"] > 0


def test_c_abi_header_is_plain_c(tmp_path):
    """include/bbfft_cuda.h is the FFI boundary: it must compile as C99 without any C++ or CUDA header."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "bbfft_cuda.h"\nint main(void) { bbfft_cuda_config c; (void)c; return bbfft_cuda_last_error() == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_empty_batch_plans_to_an_empty_grid(pkg):
    """K = 0 is a valid (empty) batch: the plan exists and launches nothing; an empty FFT shape is refused."""
    for shape, ttype, d in (([16, 64, 0], pkg.C2C, pkg.FORWARD), ([1, 8, 0], pkg.C2C, pkg.BACKWARD), ([16, 30, 0], pkg.R2C, pkg.FORWARD)):
        desc = pkg.describe(pkg.make_config(1, shape, 4, d, ttype, inplace=False))
        assert desc["grid"] == 0
    import pytest
    for shape in ([0, 64, 4], [16, 0, 4]):
        with pytest.raises(pkg.BadConfiguration):
            pkg.describe(pkg.make_config(1, shape, 4, pkg.FORWARD, pkg.C2C, inplace=False))


def test_tensor_indexer_header(tmp_path):
    """include/bbfft/tensor_indexer.hpp against the reference's expectations (test/tensor.cpp:25-155),
    compiled as plain host C++ (tests/cpp/test_tensor_indexer.cpp)."""
    import subprocess
    exe = str(tmp_path / "tti")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_tensor_indexer.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout


def test_real_nd_plans_fuse_modes_one_and_two(pkg, monkeypatch):
    """r2c / c2r 2d with an even N1 whose half-length tile fits shared memory plan as ONE fused real tile kernel
    (bbk::fft2d_tile_real_cta); odd N1, tiny tiles and BBFFT_CUDA_ND_FUSE_REAL=0 keep one launch per mode."""
    for fp in (4, 8):
        for inplace in (False, True):
            d = pkg.describe(pkg.make_config(2, [1, 128, 64, 9], fp, pkg.FORWARD, pkg.R2C, inplace=inplace))
            assert d["identifier"].startswith("bbfft_r2c2di_" if inplace else "bbfft_r2c2d_"), d["identifier"]
            # tile: N1/2 columns x N2 rows, plus the scratch column of the r2c unpack
            assert d["smem_bytes"] >= (64 * 64 + 64) * 2 * fp and d["smem_bytes"] < 2 * 64 * 64 * 2 * fp
            assert d["grid"] == 9
            d = pkg.describe(pkg.make_config(2, [1, 128, 64, 9], fp, pkg.BACKWARD, pkg.C2R, inplace=inplace))
            assert d["identifier"].startswith("bbfft_c2r2di_" if inplace else "bbfft_c2r2d_"), d["identifier"]
    for shape in ([1, 127, 64, 3], [1, 8, 8, 3]):  # odd N1; a tile below 1024 elements
        with pytest.raises(pkg.BadConfiguration):
            pkg.describe(pkg.make_config(2, shape, 4, pkg.FORWARD, pkg.R2C, inplace=False))
    monkeypatch.setenv("BBFFT_CUDA_ND_FUSE_REAL", "0")
    with pytest.raises(pkg.BadConfiguration):
        pkg.describe(pkg.make_config(2, [1, 128, 64, 9], 4, pkg.FORWARD, pkg.R2C, inplace=False))


def test_tile_kernel_variants_are_planner_decisions(pkg, monkeypatch):
    """The staged persistent tile kernel (staging buffer + bulk copies) is the default exactly for tiles that run one
    CTA per SM with more than one tile per CTA; SG / PS / BK tune keys and the environment switches override."""
    big = pkg.make_config(2, [1, 128, 128, 8192], 4, pkg.FORWARD, pkg.C2C, inplace=False)
    few = pkg.make_config(2, [1, 128, 128, 64], 4, pkg.FORWARD, pkg.C2C, inplace=False)
    small = pkg.make_config(2, [1, 64, 64, 8192], 4, pkg.FORWARD, pkg.C2C, inplace=False)
    d = pkg.describe(big)
    assert "_sg" in d["identifier"] and d["identifier"].endswith("b") and d["smem_bytes"] <= 227 * 1024
    assert "#include" in d["source"] and "STG = " in d["source"] and "BULK = 1" in d["source"]
    for cfg, tune in ((few, ""), (small, ""), (big, "SG=0")):
        ident = pkg.describe(cfg, tune)["identifier"]
        assert "_sg" not in ident and "_ps" not in ident, ident
    assert pkg.describe(big, "SG=0,PS=1")["identifier"].endswith("_ps")
    assert pkg.describe(big, "SG=32,BK=0")["identifier"].endswith("_sg4096")
    monkeypatch.setenv("BBFFT_CUDA_TILE_STAGE", "0")
    assert "_sg" not in pkg.describe(big)["identifier"]


def test_packed_add_flag_is_per_kernel(pkg, monkeypatch):
    """X2=1 (csrc/wisdom.inc, profiles/r02y_x2sweep.log): fp32 kernels of the measured sizes define BBK_F32X2 for their
    translation unit and carry _x2 in the identifier; fp64 and untuned plans never do."""
    d = pkg.describe(pkg.make_config(1, [16, 343, 64], 4, inplace=False))
    assert d["identifier"].endswith("_x2") and d["source"].lstrip().splitlines()[1].startswith("#define BBK_F32X2")
    assert not pkg.describe(pkg.make_config(1, [16, 343, 64], 8, inplace=False))["identifier"].endswith("_x2")
    assert not pkg.describe(pkg.make_config(1, [16, 512, 64], 4, inplace=False))["identifier"].endswith("_x2")
    assert not pkg.describe(pkg.make_config(1, [16, 343, 64], 4, inplace=False), "X2=0")["identifier"].endswith("_x2")
    assert pkg.describe(pkg.make_config(1, [16, 243, 64], 4, pkg.FORWARD, pkg.R2C, inplace=False))["identifier"].endswith("_x2")
    assert pkg.describe(pkg.make_config(1, [16, 343, 64], 8, inplace=False), "X2=1")["identifier"].endswith("_x2") is False
    monkeypatch.setenv("BBFFT_CUDA_NO_WISDOM", "1")
    assert not pkg.describe(pkg.make_config(1, [16, 343, 64], 4, inplace=False))["identifier"].endswith("_x2")
