"""GPU tier: 2d / 3d transforms (BASELINE config 4 and the reference's test/r2c.cpp nd cases)
through the C ABI, against numpy's float64 fftn / rfftn / irfftn on the same tensors."""
import numpy as np
import pytest

from common import TOL, cdtype, rdtype, rel_l2

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _axes(dim):
    # tensor is M x N1 x .. x Nd x K column-major = numpy (K, Nd, .., N1, M); FFT modes, N1 last
    return tuple(range(1, dim + 1))


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,Ns,K", [(1, (8, 4), 3), (3, (5, 4), 7), (2, (16, 27), 2), (1, (4, 8, 2), 3), (3, (6, 5, 4), 2),
                                    (1, (128, 128), 4), (1, (64, 64, 64), 1), (7, (10, 3), 5)])
def test_c2c_nd(pkg, fp, M, Ns, K):
    dim = len(Ns)
    rng = np.random.default_rng(sum(Ns) + M)
    shape_np = (K,) + tuple(reversed(Ns)) + (M,)
    x = (rng.standard_normal(shape_np) + 1j * rng.standard_normal(shape_np)).astype(cdtype(fp))
    for d, inplace in ((pkg.FORWARD, False), (pkg.BACKWARD, True)):
        cfg = pkg.make_config(dim, [M] + list(Ns) + [K], fp, d, pkg.C2C, inplace=inplace)
        plan = pkg.Plan(cfg, stream=_stream())
        # modes 1 and 2 run fused in one launch when their tile fits into shared memory
        fused = plan.kernel_names[0].startswith("bbfft_c2c2d")
        assert len(plan.kernel_names) == (dim - 1 if fused else dim)
        tile = M * Ns[0] * Ns[1]
        assert fused == (tile >= 1024 and tile * 2 * fp <= 200 * 1024)
        xd = torch.from_numpy(x).cuda()
        if inplace:
            plan.execute(xd)
            yd = xd
        else:
            yd = torch.empty_like(xd)
            plan.execute(xd, yd)
        torch.cuda.synchronize()
        x64 = x.astype(np.complex128)
        ref = np.fft.fftn(x64, axes=_axes(dim)) if d == pkg.FORWARD else np.fft.ifftn(x64, axes=_axes(dim)) * np.prod(Ns)
        assert rel_l2(yd.cpu().numpy(), ref) < TOL[fp], plan.kernel_names
        plan.close()


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("inplace", [False, True])
@pytest.mark.parametrize("M,Ns,K", [(1, (4, 8), 1), (3, (8, 5), 33), (1, (4, 8, 2), 33), (3, (8, 256, 5), 1), (7, (5, 4), 65),
                                    (1, (10, 3), 65), (7, (5, 4, 6), 1), (1, (10, 286, 3), 2)])
def test_r2c_c2r_nd(pkg, fp, inplace, M, Ns, K):
    """Reference test/r2c.cpp:192-252 (r2c 2d/3d in- and out-of-place) and the c2r mirrors."""
    dim = len(Ns)
    N1 = Ns[0]
    n1s = N1 // 2 + 1
    n1r = 2 * n1s if inplace else N1
    rng = np.random.default_rng(sum(Ns) + M + K)
    x = rng.uniform(-1, 1, (K,) + tuple(reversed(Ns)) + (M,)).astype(rdtype(fp))
    # ---- forward
    xin = np.zeros((K,) + tuple(reversed(Ns[1:])) + (n1r, M), dtype=rdtype(fp))
    xin[..., :N1, :] = x
    cfg = pkg.make_config(dim, [M] + list(Ns) + [K], fp, pkg.FORWARD, pkg.R2C, inplace=inplace)
    plan = pkg.Plan(cfg, stream=_stream())
    spec_shape = (K,) + tuple(reversed(Ns[1:])) + (n1s, M)
    if inplace:
        buf = torch.from_numpy(xin).cuda()
        plan.execute(buf)
        torch.cuda.synchronize()
        spec = buf.cpu().numpy().reshape(-1).view(cdtype(fp)).reshape(spec_shape)
    else:
        xd = torch.from_numpy(xin).cuda()
        yd = torch.zeros(spec_shape, dtype=torch.complex64 if fp == 4 else torch.complex128, device="cuda")
        plan.execute(xd, yd)
        torch.cuda.synchronize()
        spec = yd.cpu().numpy()
    # numpy: the halved axis is the last one in `axes` -> N1 (numpy axis dim)
    axes = tuple(range(1, dim + 1))
    ref = np.fft.rfftn(x.astype(np.float64), axes=axes)
    assert rel_l2(spec, ref) < TOL[fp], plan.kernel_names
    plan.close()
    # ---- backward (c2r) from the exact spectrum
    cfg = pkg.make_config(dim, [M] + list(Ns) + [K], fp, pkg.BACKWARD, pkg.C2R, inplace=inplace)
    plan = pkg.Plan(cfg, stream=_stream())
    sp = ref.astype(cdtype(fp))
    if inplace:
        raw = np.zeros(xin.nbytes, dtype=np.uint8)
        raw[: sp.nbytes] = np.ascontiguousarray(sp).view(np.uint8).reshape(-1)
        buf = torch.from_numpy(raw).cuda()
        plan.execute(buf)
        torch.cuda.synchronize()
        back = buf.cpu().numpy().reshape(-1).view(rdtype(fp)).reshape(xin.shape)[..., :N1, :]
    else:
        sd = torch.from_numpy(np.ascontiguousarray(sp)).cuda()
        od = torch.zeros(x.shape, dtype=torch.float32 if fp == 4 else torch.float64, device="cuda")
        plan.execute(sd, od)
        torch.cuda.synchronize()
        back = od.cpu().numpy()
    assert rel_l2(back, x.astype(np.float64) * np.prod(Ns)) < TOL[fp], plan.kernel_names
    plan.close()


def test_nd_rejects_custom_strides(pkg):
    # reference src/common/algorithm/nd_fft.hpp:47-49
    cfg = pkg.make_config(2, [2, 8, 8, 2], 4, pkg.FORWARD, pkg.C2C, istride=[1, 3, 24, 200], ostride=[1, 3, 24, 200])
    with pytest.raises(pkg.BadConfiguration):
        pkg.Plan(cfg, stream=_stream())


@pytest.mark.parametrize("fp,M,Ns,K", [(4, 1, (128, 128), 5), (8, 1, (64, 64, 64), 3), (4, 2, (32, 48), 7),
                                       (8, 1, (30, 36, 20), 4), (4, 16, (8, 8, 8), 6)])
def test_c2c_nd_fused_equals_multipass_and_l2_blocking(pkg, monkeypatch, fp, M, Ns, K):
    """BASELINE config 4 shapes: the shared-memory-fused plan, the multi-pass plan
    (BBFFT_CUDA_ND_FUSE=0, the reference's nd_fft decomposition) and L2-blocked execution
    (several k blocks per execute) agree with each other and with numpy."""
    dim = len(Ns)
    rng = np.random.default_rng(17 + sum(Ns))
    shape_np = (K,) + tuple(reversed(Ns)) + (M,)
    x = (rng.standard_normal(shape_np) + 1j * rng.standard_normal(shape_np)).astype(cdtype(fp))
    ref = np.fft.fftn(x.astype(np.complex128), axes=_axes(dim))
    xd = torch.from_numpy(x).cuda()
    per_k = x.nbytes // K
    results = {}
    for name, env in (("fused", {}), ("multipass", {"BBFFT_CUDA_ND_FUSE": "0"}),
                      ("fused-blocked", {"BBFFT_CUDA_ND_BLOCK_BYTES": str(2 * per_k)}),
                      ("multipass-unblocked", {"BBFFT_CUDA_ND_FUSE": "0", "BBFFT_CUDA_ND_BLOCK_BYTES": "0"})):
        for k in ("BBFFT_CUDA_ND_FUSE", "BBFFT_CUDA_ND_BLOCK_BYTES"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        cfg = pkg.make_config(dim, [M] + list(Ns) + [K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
        plan = pkg.Plan(cfg, stream=_stream())
        yd = torch.zeros_like(xd)
        plan.execute(xd, yd)
        torch.cuda.synchronize()
        results[name] = (yd.cpu().numpy(), plan.kernel_names, plan.launches_per_execute)
        plan.close()
    assert results["fused"][1][0].startswith("bbfft_c2c2d")
    assert not results["multipass"][1][0].startswith("bbfft_c2c2d")
    assert len(results["multipass"][1]) == dim and results["multipass-unblocked"][2] == dim
    assert results["fused-blocked"][2] == (dim - 1) * ((K + 1) // 2)
    for name, (y, names, _) in results.items():
        assert rel_l2(y, ref) < TOL[fp], (name, names)
    # blocking only changes the launch schedule, never the arithmetic
    assert np.array_equal(results["fused"][0], results["fused-blocked"][0])
    assert np.array_equal(results["multipass"][0], results["multipass-unblocked"][0])
