"""GPU tier (pytest -m gpu, B200): parity of the CUDA path, called through the C ABI
(include/bbfft_cuda.h via ctypes), against the oracle on the same seeded inputs, against the
reference's golden vectors, and through size-independent properties at the BASELINE sizes.
Tolerances are BASELINE.json's: relative L2 <= 1e-5 (fp32), 1e-12 (fp64)."""
import os

import numpy as np
import pytest

from common import TOL, analytic_c2c_input, cdtype, random_complex, reference_tol, rel_l2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _exec(pkg, cfg, x_np, out_np=None, tune=""):
    """Run one plan on numpy data via device tensors; out_np=None -> in-place.  The plan is executed
    twice (the second time into a 0xFF-filled output) and must reproduce itself bit for bit."""
    plan = pkg.Plan(cfg, stream=_stream(), tune=tune)
    xd = torch.from_numpy(x_np).cuda()
    if out_np is None:
        x2 = xd.clone()
        plan.execute(xd)
        plan.execute(x2)
        torch.cuda.synchronize()
        res = xd.cpu().numpy()
        again = x2.cpu().numpy()
        same = np.array_equal(res.view(np.uint8), again.view(np.uint8))
    else:
        yd = torch.from_numpy(out_np).cuda()
        y2 = torch.full((out_np.nbytes,), 0xFF, dtype=torch.uint8, device="cuda")
        plan.execute(xd, yd)
        plan.execute(xd, y2)
        torch.cuda.synchronize()
        res = yd.cpu().numpy()
        a, b = res.view(np.uint8).ravel(), y2.cpu().numpy()
        same = bool(np.all((a == b) | (b == 0xFF)))  # unaddressed bytes keep the fill
    names = plan.kernel_names
    plan.close()
    assert same, ("second execution differs", names)
    return res, names


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,N,K", [
    (1, 64, 16384),  # BASELINE config 1
    (16, 2, 33), (16, 3, 64), (16, 7, 100), (16, 16, 65), (16, 30, 31), (16, 64, 130), (16, 105, 17), (16, 128, 9),
    (16, 243, 5), (16, 256, 7), (16, 343, 3), (16, 500, 3), (16, 512, 5), (1, 8, 1000), (1, 15, 333), (1, 256, 37),
    (2, 12, 99), (3, 30, 50), (5, 64, 13), (17, 27, 12), (32, 25, 8), (64, 49, 3), (1024, 4, 2), (3, 363, 2), (7, 1, 3),
])
def test_c2c_vs_oracle(pkg, oracle, fp, M, N, K):
    rng = np.random.default_rng(M * 131 + N * 17 + K)
    x = random_complex(rng, (K, N, M), fp)
    d = -1 if (M + N + K) % 2 else 1
    cfg = pkg.make_config(1, [M, N, K], fp, d, pkg.C2C, inplace=False)
    y, names = _exec(pkg, cfg, x, np.zeros_like(x))
    kc = min(K, 96)  # the oracle is a long-double DFT: check a bounded prefix + suffix with it
    ocfg = oracle.make_config(1, [M, N, kc], fp, d, 0, inplace=False)
    for sl in (slice(0, kc), slice(K - kc, K)):
        ref = np.empty_like(x[sl])
        oracle.dft(ocfg, np.ascontiguousarray(x[sl]), ref)
        assert rel_l2(y[sl], ref) < TOL[fp], names
    # and the whole batch against numpy's double-precision FFT
    full = np.fft.fft(x.astype(np.complex128), axis=1) if d < 0 else np.fft.ifft(x.astype(np.complex128), axis=1) * N
    assert rel_l2(y, full) < TOL[fp], names


def test_c2c_all_sweep_sizes_both_precisions(pkg):
    """Every (precision, N) plan of the benchmark sweep is correct (vs float64 numpy)."""
    import importlib
    aot = importlib.import_module("double-batched-fft-library_b200.aot")
    rng = np.random.default_rng(5)
    for fp in (4, 8):
        for N in aot.smooth_sizes():
            K = 9
            x = random_complex(rng, (K, N, 16), fp)
            # same K-independent kernel as the 1 GiB benchmark plan
            cfg = pkg.make_config(1, [16, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
            y, names = _exec(pkg, cfg, x, np.zeros_like(x))
            ref = np.fft.fft(x.astype(np.complex128), axis=1)
            assert rel_l2(y, ref) < TOL[fp], (fp, N, names)


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M", [1, 2, 3, 16, 17, 64, 256, 1024])
@pytest.mark.parametrize("N", [2, 3, 5, 7, 11, 13, 4, 8, 16, 32, 128, 256, 512, 27, 63, 105, 363])
def test_c2c_analytic_reference_suite(pkg, fp, M, N):
    """Port of the reference's device test (test/c2c.cpp:17-66): single Fourier mode -> scaled
    delta, in-place, K in {1, 32}, per-component tolerance 1e2*eps*sqrt(N) (test/fft.hpp:17-19)."""
    for K in (1, 32):
        x, X = analytic_c2c_input(M, N, K, fp)
        cfg = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C)
        y, names = _exec(pkg, cfg, x)
        tol = reference_tol(N, fp)
        assert np.max(np.abs(y.real - X.real)) <= tol and np.max(np.abs(y.imag - X.imag)) <= tol, names


@pytest.mark.parametrize("fp", [4, 8])
def test_c2c_nonpacked_strides(pkg, oracle, fp):
    # reference test/c2c.cpp:68-82
    M, N, K = 5, 16, 33
    s = [1, M + 1, (M + 1) * (N + 1)]
    rng = np.random.default_rng(11)
    x = random_complex(rng, (K * s[2],), fp)
    cfg = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, istride=s, ostride=s)
    ocfg = oracle.make_config(1, [M, N, K], fp, -1, 0, istride=s, ostride=s)
    ref = np.zeros_like(x)
    oracle.dft(ocfg, x, ref)
    y, _ = _exec(pkg, cfg, x, np.zeros_like(x))
    assert rel_l2(y, ref) < TOL[fp]


@pytest.mark.parametrize("fp,N", [(4, 64), (4, 105), (4, 512), (8, 64), (8, 343), (8, 512)])
def test_c2c_full_size_properties(pkg, fp, N):
    """BASELINE config 2 at full size (M=16, ~1 GiB): forward o backward = N * identity
    (reference test/c2c.cpp:84-127) and linearity, both size-independent checks."""
    M = 16
    K = (1 << 30) // (M * N * 2 * fp)
    rdt = torch.float32 if fp == 4 else torch.float64
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    x = torch.view_as_complex(torch.rand(K, N, M, 2, dtype=rdt, device="cuda", generator=g))
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    fwd = pkg.Plan(pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=False), stream=_stream())
    bwd = pkg.Plan(pkg.make_config(1, [M, N, K], fp, pkg.BACKWARD, pkg.C2C, inplace=False), stream=_stream())
    fwd.execute(x, y)
    bwd.execute(y, z)
    torch.cuda.synchronize()
    err = float((z / N - x).norm() / x.norm())
    assert err < TOL[fp]
    # spot-check the spectrum of a few k slices against torch's FFT in double
    for k in (0, K // 2, K - 1):
        ref = torch.fft.fft(x[k].to(torch.complex128), dim=0)
        assert float((y[k].to(torch.complex128) - ref).norm() / ref.norm()) < TOL[fp]
    # linearity: F(a x + y) = a F(x) + F(y), in place
    a = 0.5
    w = (a * x + y).contiguous()
    fwd.execute(w, z)
    bwd.execute(y, x)  # x <- N * original x; F(y) below uses the forward plan again
    torch.cuda.synchronize()
    fy = torch.empty_like(y)
    fwd.execute(y, fy)
    torch.cuda.synchronize()
    lin = a * y + fy
    assert float((z - lin).norm() / lin.norm()) < TOL[fp]
    fwd.close()
    bwd.close()


def _golden_c2c():
    if not os.path.exists(GOLDEN):
        return []
    z = np.load(GOLDEN)
    return [str(n) for n in z["names"] if str(n).startswith("c2c")]


@pytest.mark.parametrize("name", _golden_c2c())
def test_c2c_matches_reference_golden(pkg, name):
    """Same inputs as the unmodified reference (emulated, tests/golden/make_golden.py)."""
    z = np.load(GOLDEN)
    ttype, fp, d, M, N, K, inplace = [int(v) for v in z[name + "__meta"][:7]]
    istride = [int(v) for v in z[name + "__meta"][7:10]]
    ostride = [int(v) for v in z[name + "__meta"][10:13]]
    cfg = pkg.make_config(1, [M, N, K], fp, d, ttype, istride=istride, ostride=ostride)
    x = z[name + "__in"]
    want = z[name + "__out"]
    if inplace:
        got, names = _exec(pkg, cfg, x.copy())
    else:
        got, names = _exec(pkg, cfg, x, np.zeros_like(want))
    assert rel_l2(got, want) < TOL[fp], names


def test_execute_host_and_cache(pkg):
    """The host-buffer entry point (H2D + kernel + D2H) and kernel sharing through the cache."""
    M, N, K = 16, 64, 50
    rng = np.random.default_rng(2)
    x = random_complex(rng, (K, N, M), 4)
    cache = pkg.Cache()
    p1 = pkg.Plan(pkg.make_config(1, [M, N, K], 4, inplace=False), stream=_stream(), cache=cache)
    p2 = pkg.Plan(pkg.make_config(1, [M, N, 2 * K], 4, inplace=False), stream=_stream(), cache=cache)
    assert p1.kernel_names == p2.kernel_names  # K is not part of the key
    assert len(cache) <= 1
    y = np.zeros_like(x)
    p1.execute_host(x, y)
    assert rel_l2(y, np.fft.fft(x.astype(np.complex128), axis=1)) < TOL[4]
    p1.close()
    p2.close()
    cache.close()


def test_bad_configuration_errors(pkg):
    # reference test/error.cpp:13-21: dim 0 and dim 4 throw bad_configuration
    for dim in (0, 4):
        c = pkg.make_config(1, [2, 8, 2], 4)
        c.dim = dim
        with pytest.raises(pkg.BadConfiguration):
            pkg.Plan(c, stream=_stream())


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,N,K", [(16, 37, 300), (16, 127, 100), (1, 101, 333), (3, 67, 50), (16, 254, 64), (16, 424, 40),
                                   (16, 509, 64), (1, 509, 100), (5, 379, 7)])
def test_c2c_large_prime_factors_vs_oracle(pkg, oracle, fp, M, N, K):
    """Any N works, like in the reference: prime factors beyond the in-register butterflies run as
    cooperative direct-DFT stages (bbk::run_stage_direct).  Checked against the long-double oracle,
    forward and backward, and in place."""
    rng = np.random.default_rng(N * 7 + M)
    x = (rng.standard_normal((K, N, M)) + 1j * rng.standard_normal((K, N, M))).astype(cdtype(fp))
    for d in (pkg.FORWARD, pkg.BACKWARD):
        cfg = pkg.make_config(1, [M, N, K], fp, d, pkg.C2C, inplace=False)
        plan = pkg.Plan(cfg, stream=torch.cuda.current_stream().cuda_stream)
        xd = torch.from_numpy(x).cuda()
        yd = torch.empty_like(xd)
        plan.execute(xd, yd)
        plan.execute(xd)
        torch.cuda.synchronize()
        ref = np.empty_like(x)
        oracle.dft(oracle.make_config(1, [M, N, K], fp, d, 0, inplace=False), x, ref)
        assert rel_l2(yd.cpu().numpy(), ref) < TOL[fp], plan.kernel_names
        assert np.array_equal(yd.cpu().numpy(), xd.cpu().numpy())
        plan.close()


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,N,K", [(16, 1024, 9), (1, 1024, 70), (16, 2048, 5), (3, 4096, 4), (1, 8192, 6), (16, 625, 7),
                                   (2, 1000, 11), (16, 729, 3), (1, 6561, 3), (4, 5120, 2)])
def test_c2c_beyond_512_single_kernel(pkg, fp, M, N, K):
    """N in (512, 8192]: still ONE kernel (three or four shared-memory stages; the planner narrows the
    batch lanes until a CTA's rows fit).  Outside the reference benchmark's N <= 512 sweep but inside
    what the reference accepts; round 1 covered it on the CPU emulator only."""
    rng = np.random.default_rng(N + M)
    x = random_complex(rng, (K, N, M), fp)
    for d in (pkg.FORWARD, pkg.BACKWARD):
        cfg = pkg.make_config(1, [M, N, K], fp, d, pkg.C2C, inplace=False)
        y, names = _exec(pkg, cfg, x, np.zeros_like(x))
        ref = np.fft.fft(x.astype(np.complex128), axis=1) if d < 0 else np.fft.ifft(x.astype(np.complex128), axis=1) * N
        assert rel_l2(y, ref) < TOL[fp], names
    cfg = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=True)
    y, names = _exec(pkg, cfg, x.copy())
    assert rel_l2(y, np.fft.fft(x.astype(np.complex128), axis=1)) < TOL[fp], names


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,N,K", [(16, 1024, 6), (1, 4096, 10), (2, 1250, 8), (1, 2187, 4)])
def test_real_beyond_512_single_kernel(pkg, fp, M, N, K):
    rng = np.random.default_rng(N * 3 + M)
    x = rng.uniform(-1, 1, (K, N, M)).astype(np.float32 if fp == 4 else np.float64)
    cfg = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.R2C, inplace=False)
    spec = np.zeros((K, N // 2 + 1, M), dtype=cdtype(fp))
    y, names = _exec(pkg, cfg, x, spec)
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    assert rel_l2(y, ref) < TOL[fp], names
    cfg = pkg.make_config(1, [M, N, K], fp, pkg.BACKWARD, pkg.C2R, inplace=False)
    back, names = _exec(pkg, cfg, y, np.zeros_like(x))
    assert rel_l2(back / N, x) < TOL[fp] * 4, names
