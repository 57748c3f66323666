"""GPU tier: the host-buffer entry point bbfft_cuda_plan_execute_host (H2D + transform + D2H behind one
call: the `e2e` leg of bench.py).  Slab pipelining over the device ring must be invisible: results are
bit-identical to the device-resident execute for every layout, and layouts whose k slices are not
contiguous byte ranges fall back to one whole-tensor copy (round-1 advisor finding)."""
import numpy as np
import pytest

from common import C2R, R2C, TOL, cdtype, rdtype, real_problem, rel_l2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _device_result(pkg, cfg, x, out_bytes, inplace):
    plan = pkg.Plan(cfg, stream=_stream())
    if inplace:
        raw = np.zeros(max(x.nbytes, out_bytes), np.uint8)
        raw[: x.nbytes] = x.view(np.uint8)
        d = torch.from_numpy(raw).cuda()
        plan.execute(d)
        res = d.cpu().numpy()
    else:
        d = torch.from_numpy(x.view(np.uint8)).cuda()
        y = torch.zeros(out_bytes, dtype=torch.uint8, device="cuda")
        plan.execute(d, y)
        res = y.cpu().numpy()
    plan.close()
    return res


@pytest.mark.parametrize("fp,M,N,K", [(4, 16, 64, 40000), (8, 16, 105, 9000), (4, 1, 256, 70000), (8, 3, 30, 100001)])
def test_c2c_ring_equals_device_path(pkg, fp, M, N, K):
    """> 32 MiB tensors: several slabs in flight on the ring; the last slab is ragged."""
    rng = np.random.default_rng(N + K)
    x = (rng.standard_normal((K, N, M)) + 1j * rng.standard_normal((K, N, M))).astype(cdtype(fp))
    cfg = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
    want = _device_result(pkg, cfg, x.reshape(-1), x.nbytes, False).view(cdtype(fp))
    plan = pkg.Plan(cfg, stream=_stream())
    hin = torch.from_numpy(x.reshape(-1).copy()).pin_memory()
    hout = torch.zeros(x.size, dtype=hin.dtype).pin_memory()
    for _ in range(2):  # a second call re-uses slots that still carry the first call's events
        hout.zero_()
        plan.execute_host(hin, hout)
        assert np.array_equal(hout.numpy(), want)
    # in place on the host buffer
    plan_ip = pkg.Plan(pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=True), stream=_stream())
    plan_ip.execute_host(hin)
    assert np.array_equal(hin.numpy(), want)
    ref = np.fft.fft(x[:64].astype(np.complex128), axis=1)
    assert rel_l2(want.reshape(K, N, M)[:64], ref) < TOL[fp]
    plan.close()
    plan_ip.close()


@pytest.mark.parametrize("ttype", [R2C, C2R])
@pytest.mark.parametrize("fp,M,N,K,inplace", [(4, 1, 256, 60001, False), (4, 1, 256, 60001, True), (8, 16, 105, 5001, False),
                                              (4, 3, 27, 200001, False), (8, 1, 30, 200000, True)])
def test_real_ring_equals_device_path(pkg, ttype, fp, M, N, K, inplace):
    """Real transforms through the ring: padded in-place rows, odd N (slice pairs 2k', 2k'+1 must stay
    in one slab), odd K."""
    d = -1 if ttype == R2C else 1
    cfg = pkg.make_config(1, [M, N, K], fp, d, ttype, inplace=inplace)
    if inplace and pkg.describe(cfg)["inplace_unsupported"]:
        pytest.skip("in-place unsupported for this shape (reference semantics)")
    rng = np.random.default_rng(7 * N + K + ttype)
    ist, ost, x, odt, nout = real_problem(rng, pkg, ttype, M, N, K, fp, inplace, pollute=(ttype == C2R))
    out_bytes = nout * np.dtype(odt).itemsize
    want = _device_result(pkg, cfg, x, out_bytes, inplace)
    plan = pkg.Plan(cfg, stream=_stream())
    if inplace:
        raw = np.zeros(max(x.nbytes, out_bytes), np.uint8)
        raw[: x.nbytes] = x.view(np.uint8)
        plan.execute_host(raw)
        got = raw
    else:
        got = np.zeros(out_bytes, np.uint8)
        plan.execute_host(x.view(np.uint8), got)
    plan.close()
    assert np.array_equal(got[: want.size], want)


def test_interleaved_k_layout_is_not_chunked(pkg):
    """istride = {1, M*K, M}: k is interleaved with n, a k slice is not one byte range.  The host path
    must copy the whole tensor (round 1 uploaded only part of it once the tensor exceeded 32 MiB)."""
    M, N, K = 16, 64, 5000
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((N, K, M)) + 1j * rng.standard_normal((N, K, M))).astype(np.complex64)  # [n][k][m]
    strides = [1, M * K, M]
    cfg = pkg.make_config(1, [M, N, K], 4, pkg.FORWARD, pkg.C2C, istride=strides, ostride=strides)
    plan = pkg.Plan(cfg, stream=_stream())
    y = np.zeros_like(x)
    assert x.nbytes > (32 << 20)
    plan.execute_host(x.reshape(-1), y.reshape(-1))
    plan.close()
    ref = np.fft.fft(x.astype(np.complex128), axis=0)
    assert rel_l2(y, ref) < TOL[4]


def test_undersized_host_buffers_are_rejected(pkg):
    M, N, K = 16, 64, 100
    cfg = pkg.make_config(1, [M, N, K], 4, pkg.FORWARD, pkg.C2C, inplace=False)
    plan = pkg.Plan(cfg, stream=_stream())
    x = np.zeros(M * N * K, np.complex64)
    with pytest.raises(pkg.BadConfiguration):
        plan.execute_host(x[:-1], np.zeros_like(x))
    with pytest.raises(pkg.BadConfiguration):
        plan.execute_host(x, np.zeros(M * N * K - 1, np.complex64))
    plan.execute_host(x, np.zeros_like(x))
    plan.close()


def test_nd_plan_through_host_path(pkg):
    dims, K = (32, 16), 300
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((K, dims[1], dims[0])) + 1j * rng.standard_normal((K, dims[1], dims[0]))).astype(np.complex64)
    cfg = pkg.make_config(2, [1, dims[0], dims[1], K], 4, pkg.FORWARD, pkg.C2C, inplace=False)
    plan = pkg.Plan(cfg, stream=_stream())
    y = np.zeros_like(x)
    plan.execute_host(x.reshape(-1), y.reshape(-1))
    plan.close()
    assert rel_l2(y, np.fft.fftn(x.astype(np.complex128), axes=(1, 2))) < TOL[4]


def test_unaligned_real_pointer_m1(pkg):
    """M = 1 even-N real transforms read the real tensor as aligned complex words; a real pointer at an
    odd element offset (a view x[1:], which the reference accepts) is served by the PAIR=0 kernel."""
    N, K = 64, 257
    for fp in (4, 8):
        rdt, cdt = rdtype(fp), cdtype(fp)
        rng = np.random.default_rng(fp)
        xr = rng.uniform(-1, 1, (K, N)).astype(rdt)
        buf = torch.zeros(K * N + 1, dtype=torch.float32 if fp == 4 else torch.float64, device="cuda")
        buf[1:] = torch.from_numpy(xr.reshape(-1)).cuda()
        view = buf[1:]
        assert view.data_ptr() % (2 * fp) != 0
        cfg = pkg.make_config(1, [1, N, K], fp, pkg.FORWARD, pkg.R2C, inplace=False)
        plan = pkg.Plan(cfg, stream=_stream())
        out = torch.zeros(K * (N // 2 + 1), dtype=torch.complex64 if fp == 4 else torch.complex128, device="cuda")
        plan.execute(view, out)
        torch.cuda.synchronize()
        ref = np.fft.rfft(xr.astype(np.float64), axis=1)
        assert rel_l2(out.cpu().numpy().reshape(K, N // 2 + 1), ref) < TOL[fp]
        # and back: c2r into an odd-offset real view
        cfgb = pkg.make_config(1, [1, N, K], fp, pkg.BACKWARD, pkg.C2R, inplace=False)
        planb = pkg.Plan(cfgb, stream=_stream())
        back = torch.zeros_like(buf)
        planb.execute(out, back[1:])
        torch.cuda.synchronize()
        assert rel_l2(back[1:].cpu().numpy().reshape(K, N) / N, xr) < TOL[fp]
        plan.close()
        planb.close()
