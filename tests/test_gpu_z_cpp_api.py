"""The C++ drop-in API exercised from compiled user code (tests/cpp/test_cuda_api.cpp): compiled
in the CPU tier, executed on the GPU tier."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "double-batched-fft-library_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_cuda_api")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "test_cuda_api.cpp")
    deps = [src, os.path.join(LIBDIR, "libbbfft_cuda.so")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(CUDA_HOME, "include"), src, "-o", EXE, "-L" + LIBDIR, "-lbbfft_cuda",
                           "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcudart_static", "-ldl", "-lrt", "-lpthread",
                           "-Wl,-rpath," + LIBDIR])


REPRO = os.path.join(ROOT, "tests", "cpp", "repro_c2r")


def _build_repro():
    src = os.path.join(ROOT, "tests", "cpp", "repro_c2r.cpp")
    deps = [src, os.path.join(LIBDIR, "libbbfft_cuda.so")]
    if os.path.exists(REPRO) and all(os.path.getmtime(REPRO) >= os.path.getmtime(d) for d in deps):
        return
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(CUDA_HOME, "include"), src, "-o", REPRO, "-L" + LIBDIR, "-lbbfft_cuda",
                           "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcudart_static", "-ldl", "-lrt", "-lpthread",
                           "-Wl,-rpath," + LIBDIR])


def test_soak_harness_compiles(pkg):
    _build_repro()
    assert os.path.exists(REPRO)


@pytest.mark.gpu
@pytest.mark.parametrize("args", [
    # the round-1 wrong result: fp64 c2r M=32 N=424 (4 x 53 half length), fresh plans and one re-used plan, user
    # streams of both kinds; every execute against a host long-double DFT and bit-identical to the first
    ["c2r", "f64", "32", "424", "4", "40", "fresh", "blocking"],
    ["c2r", "f64", "32", "424", "4", "300", "reuse", "nonblocking"],
    ["c2r", "f32", "32", "424", "4", "100", "reuse", "blocking"],
    ["r2c", "f64", "32", "848", "4", "100", "reuse", "nonblocking"],
    ["c2c", "f64", "16", "490", "16", "100", "reuse", "default"],
    ["c2c", "f32", "16", "509", "8", "50", "reuse", "blocking"],
])
def test_determinism_soak(pkg, args):
    """tests/cpp/repro_c2r: repeated executes from C++ user code (the environment the round-1 failure needed)."""
    _build_repro()
    r = subprocess.run([REPRO] + args, capture_output=True, text=True, timeout=900)
    print(r.stdout[-2000:])
    assert r.returncode == 0 and "0 bad iterations" in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]


def test_cpp_api_user_code_compiles(pkg):
    """User code written against the reference's headers (configuration aggregate init,
    make_plan, execute overloads, caches, generator) builds against include/bbfft."""
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_api_on_device(pkg):
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=1500)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_empty_batch_executes_as_a_no_op(pkg):
    """K = 0: plan creation succeeds, execute launches nothing and leaves the output untouched."""
    cfg = pkg.make_config(1, [16, 64, 0], 4, pkg.FORWARD, pkg.C2C, inplace=False)
    import torch
    plan = pkg.Plan(cfg, stream=torch.cuda.current_stream().cuda_stream)
    x = torch.zeros(8, dtype=torch.complex64, device="cuda")
    y = torch.full((8,), 3.0, dtype=torch.complex64, device="cuda")
    plan.execute(x, y)
    torch.cuda.synchronize()
    assert bool((y == 3.0).all())
    plan.close()


@pytest.mark.gpu
def test_execute_is_capturable_in_a_cuda_graph(pkg):
    """plan::execute is one stream-ordered kernel launch with by-value arguments, so a launch-bound
    loop (BASELINE config 1: 8 MiB per transform batch, ~6 us per launch) can be captured in a CUDA
    graph and replayed."""
    import torch
    cfg = pkg.make_config(1, [1, 64, 16384], 4, pkg.FORWARD, pkg.C2C, inplace=False)
    plan = pkg.Plan(cfg, stream=torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    x = torch.view_as_complex(torch.randn(16384 * 64, 2, device="cuda", generator=gen))
    want = torch.empty_like(x)
    plan.execute(x, want)
    torch.cuda.synchronize()
    y = torch.zeros_like(x)
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(8):
            plan.execute(x, y, stream=torch.cuda.current_stream().cuda_stream)
    y.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y, want)
    plan.close()
