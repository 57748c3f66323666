"""GPU tier: user load/store callbacks fused into the FFT kernel with NVRTC.  Port of the
reference's test/callback.cpp (:18-114 load callback, :116-202 store callback): results must be
BIT-IDENTICAL (==) to a plain plan on the zero-padded / untruncated problem."""
import numpy as np
import pytest

from callbacks import load_zero_pad_cuda, load_zero_pad_opencl, store_truncate_scale_opencl
from common import TOL, rel_l2, cdtype, rdtype

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _run(pkg, cfg, x, out):
    plan = pkg.Plan(cfg, stream=_stream())
    xd = torch.from_numpy(x).cuda()
    yd = torch.from_numpy(out).cuda()
    plan.execute(xd, yd)
    torch.cuda.synchronize()
    names = plan.kernel_names
    plan.close()
    return yd.cpu().numpy(), names


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("lang", ["opencl", "cuda"])
@pytest.mark.parametrize("M", [1, 32])
@pytest.mark.parametrize("N", [8, 64, 212])
def test_load_callback(pkg, fp, lang, M, N):
    K = 16
    real = "float" if fp == 4 else "double"
    rng = np.random.default_rng(N * 10 + M)
    N_ext = 2 * N
    ns_ref, ns = N_ext // 2 + 1, N // 2 + 1
    X = (rng.uniform(0, 1, (K, ns, M)) + 1j * rng.uniform(0, 1, (K, ns, M))).astype(cdtype(fp))
    X_ref = np.zeros((K, ns_ref, M), dtype=cdtype(fp))
    X_ref[:, :ns, :] = X
    strides = dict(istride=[1, M, M * ns_ref], ostride=[1, M, M * N_ext])
    src = (load_zero_pad_opencl if lang == "opencl" else load_zero_pad_cuda)(real, M, ns_ref, ns)
    cfg_ref = pkg.make_config(1, [M, N_ext, K], fp, pkg.BACKWARD, pkg.C2R, **strides)
    cfg = pkg.make_config(1, [M, N_ext, K], fp, pkg.BACKWARD, pkg.C2R, callbacks=(src, "load", None, lang), **strides)
    x_ref, ref_names = _run(pkg, cfg_ref, X_ref.reshape(-1), np.zeros(K * N_ext * M, dtype=rdtype(fp)))
    x, names = _run(pkg, cfg, X.reshape(-1), np.zeros(K * N_ext * M, dtype=rdtype(fp)))
    assert names[0].endswith("_load")
    assert np.array_equal(x, x_ref), names
    # two plans that are wrong the same way would pass the == above (round 1: fp64 M=32 N=212 was):
    # the plain plan is also held against an independent fp64 transform, and must reproduce itself
    Xh = X_ref.astype(np.complex128)
    Xh[:, 0, :] = Xh[:, 0, :].real          # imag(X[0]) is ignored by c2r (test/r2c.cpp:310-324)
    Xh[:, ns_ref - 1, :] = Xh[:, ns_ref - 1, :].real  # zero-padded: the Nyquist row is 0 anyway
    want = np.fft.irfft(Xh, n=N_ext, axis=1) * N_ext
    assert rel_l2(x_ref.reshape(K, N_ext, M), want) < TOL[fp], ref_names
    again, _ = _run(pkg, cfg_ref, X_ref.reshape(-1), np.full(K * N_ext * M, np.nan, dtype=rdtype(fp)))
    assert np.array_equal(again, x_ref), ref_names


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M", [1, 32])
@pytest.mark.parametrize("N", [8, 64, 212])
def test_store_callback(pkg, fp, M, N):
    K = 16
    real = "float" if fp == 4 else "double"
    rng = np.random.default_rng(N * 10 + M + 1)
    ncut, nspec = N // 4, N // 2 + 1
    xin = rng.uniform(0, 1, K * N * M).astype(rdtype(fp))
    strides = dict(istride=[1, M, M * N], ostride=[1, M, M * nspec])
    cfg_ref = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.R2C, **strides)
    cfg = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.R2C,
                          callbacks=(store_truncate_scale_opencl(real, M, nspec, ncut, 1.0 / N), None, "store", "opencl"),
                          **strides)
    Y_ref, _ = _run(pkg, cfg_ref, xin, np.zeros(K * nspec * M, dtype=cdtype(fp)))
    Y, names = _run(pkg, cfg, xin, np.zeros(K * ncut * M, dtype=cdtype(fp)))
    want = (Y_ref.reshape(K, nspec, M)[:, :ncut, :] * rdtype(fp)(1.0 / N)).astype(cdtype(fp))
    assert names[0].endswith("_store")
    assert np.array_equal(Y.reshape(K, ncut, M), want), names
    full = np.fft.rfft(xin.astype(np.float64).reshape(K, N, M), axis=1)
    assert rel_l2(Y_ref.reshape(K, nspec, M), full) < TOL[fp]


@pytest.mark.parametrize("fp,N", [(4, 64), (4, 256), (8, 64)])
def test_config5_identity_callbacks_full_size(pkg, fp, N):
    """BASELINE config 5: double-batched c2c (M=16, K as config 2) with identity load/store
    callbacks: bit-identical to the plain plan (which comes from the nvcc-built bundle, so this
    also pins nvcc-AOT == NVRTC-JIT arithmetic)."""
    M = 16
    K = (1 << 28) // (M * N * 2 * fp)
    v = "float2" if fp == 4 else "double2"
    src = ("__device__ %s my_load(%s const* in, size_t off) { return in[off]; }\n"
           "__device__ void my_store(%s* out, size_t off, %s v) { out[off] = v; }\n") % (v, v, v, v)
    rdt = torch.float32 if fp == 4 else torch.float64
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    x = torch.view_as_complex(torch.rand(K, N, M, 2, dtype=rdt, device="cuda", generator=g))
    y0 = torch.empty_like(x)
    y1 = torch.empty_like(x)
    p0 = pkg.Plan(pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=False), stream=_stream())
    p1 = pkg.Plan(pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=False,
                                  callbacks=(src, "my_load", "my_store", "cuda")), stream=_stream())
    p0.execute(x, y0)
    p1.execute(x, y1)
    torch.cuda.synchronize()
    assert p1.kernel_names[0].endswith("_my_load_my_store")
    assert torch.equal(torch.view_as_real(y0), torch.view_as_real(y1))
    p0.close()
    p1.close()


def test_callbacks_rejected_for_nd(pkg):
    # reference src/common/algorithm/nd_fft.hpp:29-31
    src = "__device__ float2 l(float2 const* in, size_t off) { return in[off]; }"
    cfg = pkg.make_config(2, [1, 8, 8, 2], 4, pkg.FORWARD, pkg.C2C, callbacks=(src, "l", None, "cuda"))
    with pytest.raises(pkg.BadConfiguration):
        pkg.Plan(cfg, stream=_stream())
