"""GPU tier: 2d / 3d transforms (BASELINE config 4 and the reference's test/r2c.cpp nd cases)
through the C ABI, against numpy's float64 fftn / rfftn / irfftn on the same tensors."""
import numpy as np
import pytest

from common import C2R, R2C, TOL, cdtype, rdtype, rel_l2

C2C = 0

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _axes(dim):
    # tensor is M x N1 x .. x Nd x K column-major = numpy (K, Nd, .., N1, M); FFT modes, N1 last
    return tuple(range(1, dim + 1))


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,Ns,K", [(1, (8, 4), 3), (3, (5, 4), 7), (2, (16, 27), 2), (1, (4, 8, 2), 3), (3, (6, 5, 4), 2),
                                    (1, (128, 128), 4), (1, (64, 64, 64), 1), (7, (10, 3), 5),
                                    (1, (256, 256), 3), (4, (64, 64), 5), (2, (128, 32), 9), (1, (96, 80), 7),
                                    (16, (16, 64), 6)])
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
    chain = {"BBFFT_CUDA_ND_CHAIN": "1"}
    for name, env in (("chained", chain), ("fused", {}), ("multipass", {"BBFFT_CUDA_ND_FUSE": "0"}),
                      ("fused-blocked", {"BBFFT_CUDA_ND_BLOCK_BYTES": str(2 * per_k)}),
                      ("multipass-unblocked", {"BBFFT_CUDA_ND_FUSE": "0", "BBFFT_CUDA_ND_BLOCK_BYTES": "0"}),
                      ("multipass-chained", dict(chain, BBFFT_CUDA_ND_FUSE="0", BBFFT_CUDA_ND_CHAIN_KBLOCK="1"))):
        for k in ("BBFFT_CUDA_ND_FUSE", "BBFFT_CUDA_ND_BLOCK_BYTES", "BBFFT_CUDA_ND_CHAIN", "BBFFT_CUDA_ND_CHAIN_KBLOCK"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        cfg = pkg.make_config(dim, [M] + list(Ns) + [K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
        plan = pkg.Plan(cfg, stream=_stream())
        yd = torch.zeros_like(xd)
        for _ in range(3):  # chained plans count launches (epochs): repeated executes must agree
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
    # one persistent launch for all steps where they can share a CTA shape (bbk::chain); it only
    # changes the schedule, never the arithmetic
    for name in ("chained", "multipass-chained"):
        if results[name][1][0].startswith("bbfft_chain"):
            assert results[name][2] == 1 and len(results[name][1]) == 1
    if dim == 3 and M == 1 and Ns[0] >= 32:
        assert results["chained"][1][0].startswith("bbfft_chain2"), results["chained"][1]
    assert np.array_equal(results["chained"][0], results["fused"][0])
    assert np.array_equal(results["multipass-chained"][0], results["multipass"][0])
    # blocking only changes the launch schedule, never the arithmetic
    assert np.array_equal(results["fused"][0], results["fused-blocked"][0])
    assert np.array_equal(results["multipass"][0], results["multipass-unblocked"][0])


@pytest.mark.parametrize("fp,ttype,Ns,K", [(4, R2C, (128, 64), 6), (4, C2R, (128, 64), 6), (8, R2C, (256, 48), 5),
                                          (4, C2C, (1024, 64), 3), (8, C2C, (512, 512), 2), (4, C2C, (256, 128, 64), 2),
                                          (4, R2C, (128, 64, 32), 2), (8, C2R, (128, 64, 64), 2)])
def test_nd_chain_matches_one_launch_per_step(pkg, monkeypatch, fp, ttype, Ns, K):
    """Chains of double-batched 1d passes (no fused tile: r2c/c2r, tiles beyond shared memory, three
    passes): same bits as one launch per step, and right against numpy."""
    dim = len(Ns)
    rng = np.random.default_rng(5 + sum(Ns))
    sig_shape = (K,) + tuple(reversed(Ns)) + (1,)
    xr = rng.standard_normal(sig_shape)
    if ttype == C2C:
        x = (xr + 1j * rng.standard_normal(sig_shape)).astype(cdtype(fp))
        ref = np.fft.fftn(x.astype(np.complex128), axes=_axes(dim))
        nout, odt, d = x.size, cdtype(fp), pkg.FORWARD
    elif ttype == R2C:
        x = xr.astype(rdtype(fp))
        ref = np.fft.fftn(np.fft.rfft(x.astype(np.float64), axis=dim), axes=_axes(dim)[:-1]) if dim > 1 else None
        nout, odt, d = ref.size, cdtype(fp), pkg.FORWARD
    else:
        spec = np.fft.fftn(np.fft.rfft(xr, axis=dim), axes=_axes(dim)[:-1])
        x = spec.astype(cdtype(fp))
        ref = xr * np.prod(Ns)
        nout, odt, d = ref.size, rdtype(fp), pkg.BACKWARD
    outs = {}
    # (real transforms: one launch per mode on both sides -- the fused real tile kernel is a different factorization)
    for name, env in (("chain", {"BBFFT_CUDA_ND_CHAIN": "1"}), ("steps", {"BBFFT_CUDA_ND_FUSE_REAL": "0"})):
        monkeypatch.delenv("BBFFT_CUDA_ND_CHAIN", raising=False)
        monkeypatch.delenv("BBFFT_CUDA_ND_FUSE_REAL", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        cfg = pkg.make_config(dim, [1] + list(Ns) + [K], fp, d, ttype, inplace=False)
        plan = pkg.Plan(cfg, stream=_stream())
        xd = torch.from_numpy(x).cuda()
        yd = torch.zeros(nout, dtype=torch.from_numpy(np.zeros(1, odt)).dtype, device="cuda")
        for _ in range(2):
            plan.execute(xd, yd)
        torch.cuda.synchronize()
        outs[name] = (yd.cpu().numpy(), plan.kernel_names, plan.launches_per_execute)
        plan.close()
    assert outs["chain"][1][0].startswith("bbfft_chain") and outs["chain"][2] == 1, outs["chain"][1]
    assert outs["steps"][2] == dim
    assert rel_l2(outs["chain"][0].reshape(ref.shape), ref) < TOL[fp]
    assert np.array_equal(outs["chain"][0], outs["steps"][0])


def test_config4_3d_chain_full_size(pkg, monkeypatch):
    """BASELINE config 4 at full size (3d c2c fp64 64^3, K=64, 256 MiB): the chained plan equals the
    two-launch plan bit for bit, in place as well, over repeated executes."""
    K = 64
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    x = torch.randn(K * 64 * 64 * 64, 2, dtype=torch.float64, device="cuda", generator=gen)
    outs = {}
    for name, env in (("chain", {"BBFFT_CUDA_ND_CHAIN": "1"}), ("steps", {})):
        monkeypatch.delenv("BBFFT_CUDA_ND_CHAIN", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        cfg = pkg.make_config(3, [1, 64, 64, 64, K], 8, pkg.FORWARD, pkg.C2C, inplace=False)
        plan = pkg.Plan(cfg, stream=_stream())
        y = torch.zeros_like(x)
        for _ in range(3):
            plan.execute(x, y)
        z = x.clone()
        plan.execute(z)
        torch.cuda.synchronize()
        outs[name] = (y, z, plan.kernel_names)
        plan.close()
    assert outs["chain"][2][0].startswith("bbfft_chain2")
    assert torch.equal(outs["chain"][0], outs["steps"][0])
    assert torch.equal(outs["chain"][1], outs["steps"][0])
    want = torch.fft.fftn(torch.view_as_complex(x).view(K, 64, 64, 64)[:2], dim=(1, 2, 3))
    got = torch.view_as_complex(outs["chain"][0]).view(K, 64, 64, 64)[:2]
    assert float((got - want).norm() / want.norm()) < TOL[8]


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("cl", [2, 4, 8])
@pytest.mark.parametrize("M,Ns,K", [(1, (128, 128), 4), (1, (64, 64), 7), (1, (256, 256), 3), (1, (512, 128), 2), (4, (64, 64), 5),
                                    (2, (128, 32), 9), (1, (64, 128, 4), 2), (1, (64, 64, 64), 1)])
def test_c2c_nd_tile_split_over_a_cluster(pkg, monkeypatch, fp, cl, M, Ns, K):
    """BBFFT_CUDA_TILE_CLUSTER=<cl>: the fused tile is split over a thread-block cluster (rows per CTA in
    the row pass, columns in the column pass, gathered through distributed shared memory); tiles beyond
    one CTA's shared memory become fusable.  Off by default (measured slower), kept correct: forward out
    of place and backward in place against numpy, and bit-identical to the one-CTA-per-tile kernel."""
    dim = len(Ns)
    rng = np.random.default_rng(sum(Ns) + M + cl)
    shape_np = (K,) + tuple(reversed(Ns)) + (M,)
    x = (rng.standard_normal(shape_np) + 1j * rng.standard_normal(shape_np)).astype(cdtype(fp))
    outs = {}
    for mode in ("cluster", "plain"):
        if mode == "cluster":
            monkeypatch.setenv("BBFFT_CUDA_TILE_CLUSTER", str(cl))
        else:
            monkeypatch.delenv("BBFFT_CUDA_TILE_CLUSTER", raising=False)
        res = []
        for d, inplace in ((pkg.FORWARD, False), (pkg.BACKWARD, True)):
            cfg = pkg.make_config(dim, [M] + list(Ns) + [K], fp, d, pkg.C2C, inplace=inplace)
            plan = pkg.Plan(cfg, stream=_stream())
            xd = torch.from_numpy(x).cuda()
            if inplace:
                plan.execute(xd)
                yd = xd
            else:
                yd = torch.empty_like(xd)
                plan.execute(xd, yd)
            torch.cuda.synchronize()
            res.append((yd.cpu().numpy(), plan.kernel_names))
            plan.close()
        outs[mode] = res
    x64 = x.astype(np.complex128)
    refs = [np.fft.fftn(x64, axes=_axes(dim)), np.fft.ifftn(x64, axes=_axes(dim)) * np.prod(Ns)]
    for (y, names), ref in zip(outs["cluster"], refs):
        assert rel_l2(y, ref) < TOL[fp], names
    tile = M * Ns[0] * Ns[1]
    if tile * 2 * fp <= 200 * 1024 and Ns[0] % cl == 0 and Ns[1] % cl == 0 and tile * 2 * fp // cl >= 16 * 1024:
        for (y, names), (y0, _) in zip(outs["cluster"], outs["plain"]):
            if "_cl" in names[0]:
                assert np.array_equal(y, y0), names


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("env", [{"BBFFT_CUDA_TILE_STAGE": "-1"}, {"BBFFT_CUDA_TILE_STAGE": "8", "BBFFT_CUDA_TILE_BULK": "0"},
                                 {"BBFFT_CUDA_TILE_STAGE": "-1", "BBFFT_CUDA_TILE_BULK": "1"},
                                 {"BBFFT_CUDA_TILE_ASYNC": "1", "BBFFT_CUDA_TILE_STAGE": "0"}])
@pytest.mark.parametrize("M,Ns,K", [(1, (128, 128), 700), (1, (64, 64), 1500), (4, (64, 64), 5), (1, (40, 30), 1000), (1, (64, 64, 64), 3)])
def test_c2c_nd_persistent_tile_pipelines(pkg, monkeypatch, fp, env, M, Ns, K):
    """The persistent tile kernels: BBFFT_CUDA_TILE_STAGE=<rows> (bbk::fft2d_tile_staged: the leading rows of the
    next tile wait in a staging buffer beside the tile) and BBFFT_CUDA_TILE_ASYNC=1 (bbk::fft2d_tile_persistent).
    K is large enough that every CTA of the resident grid walks several tiles.  Forward out of place and
    backward in place against numpy, bit-identical to the one-tile-per-CTA kernel, and the same twice."""
    dim = len(Ns)
    rng = np.random.default_rng(sum(Ns) + M + K)
    shape_np = (K,) + tuple(reversed(Ns)) + (M,)
    x = (rng.standard_normal(shape_np) + 1j * rng.standard_normal(shape_np)).astype(cdtype(fp))
    outs = {}
    for mode in ("pipe", "plain"):
        for k in ("BBFFT_CUDA_TILE_STAGE", "BBFFT_CUDA_TILE_ASYNC", "BBFFT_CUDA_TILE_BULK"):
            monkeypatch.delenv(k, raising=False)
        if mode == "pipe":
            for k, v in env.items():
                monkeypatch.setenv(k, v)
        else:
            monkeypatch.setenv("BBFFT_CUDA_TILE_STAGE", "0")  # (staging is the default for one-CTA-per-SM tiles)
        res = []
        for d, inplace in ((pkg.FORWARD, False), (pkg.BACKWARD, True)):
            cfg = pkg.make_config(dim, [M] + list(Ns) + [K], fp, d, pkg.C2C, inplace=inplace)
            plan = pkg.Plan(cfg, stream=_stream())
            xd = torch.from_numpy(x).cuda()
            if inplace:
                plan.execute(xd)
                yd = xd
            else:
                yd = torch.empty_like(xd)
                plan.execute(xd, yd)
                y1 = yd.clone()
                yd.zero_()
                plan.execute(xd, yd)
                assert torch.equal(y1, yd), plan.kernel_names
            torch.cuda.synchronize()
            res.append((yd.cpu().numpy(), plan.kernel_names))
            plan.close()
        outs[mode] = res
    tag = "_ps" if env.get("BBFFT_CUDA_TILE_ASYNC") == "1" else "_sg"
    if outs["plain"][0][1][0].startswith("bbfft_c2c2d"):  # (tiles beyond one CTA's shared memory run one pass per mode)
        assert tag in outs["pipe"][0][1][0], outs["pipe"][0][1]
    assert tag not in outs["plain"][0][1][0]
    x64 = x[: min(K, 3)].astype(np.complex128)
    refs = [np.fft.fftn(x64, axes=_axes(dim)), np.fft.ifftn(x64, axes=_axes(dim)) * np.prod(Ns)]
    for (y, names), (y0, _), ref in zip(outs["pipe"], outs["plain"], refs):
        assert rel_l2(y[: min(K, 3)], ref) < TOL[fp], names
        assert np.array_equal(y, y0), names


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("inplace", [False, True])
@pytest.mark.parametrize("M,Ns,K", [(1, (128, 128), 40), (1, (64, 64), 300), (3, (20, 36), 50), (2, (12, 160), 33), (16, (8, 16), 20),
                                    (1, (128, 16), 9), (1, (40, 60), 64), (1, (256, 64), 7), (1, (64, 64, 8), 5), (1, (64, 32, 32), 4),
                                    (4, (16, 32, 3), 6)])
def test_real_nd_fused_tiles(pkg, monkeypatch, fp, inplace, M, Ns, K):
    """Fused real tile kernels (bbk::fft2d_tile_real_cta: modes 1 and 2 of an r2c / c2r transform in one launch,
    reference src/common/algorithm/nd_fft.hpp:66-152 runs one launch per mode): against numpy's float64 rfftn /
    irfftn, against the one-launch-per-mode plan (BBFFT_CUDA_ND_FUSE_REAL=0), the same bits on a second execute,
    and -- out of place, M = 1 -- with a real tensor that is not aligned to a complex number (falls back)."""
    dim = len(Ns)
    N1 = Ns[0]
    n1s = N1 // 2 + 1
    n1r = 2 * n1s if inplace else N1
    rng = np.random.default_rng(sum(Ns) + M + K + 3)
    x = rng.uniform(-1, 1, (K,) + tuple(reversed(Ns)) + (M,)).astype(rdtype(fp))
    axes = tuple(range(1, dim + 1))
    ref = np.fft.rfftn(x.astype(np.float64), axes=axes)
    spec_shape = (K,) + tuple(reversed(Ns[1:])) + (n1s, M)
    rt = torch.float32 if fp == 4 else torch.float64
    ct = torch.complex64 if fp == 4 else torch.complex128
    xin = np.zeros((K,) + tuple(reversed(Ns[1:])) + (n1r, M), dtype=rdtype(fp))
    xin[..., :N1, :] = x
    sp = np.ascontiguousarray(ref.astype(cdtype(fp)))
    sp.reshape(K, -1)[:, 0] += 0.37j  # the imaginary part of X[0] must be ignored (reference test/r2c.cpp:310-324)
    got = {}
    # (the one-launch-per-mode twin compiles four more kernels per case: fp32 out of place only)
    modes = ("fused", "unfused") if (fp == 4 and not inplace) else ("fused",)
    for mode in modes:
        monkeypatch.delenv("BBFFT_CUDA_ND_FUSE_REAL", raising=False)
        if mode == "unfused":
            monkeypatch.setenv("BBFFT_CUDA_ND_FUSE_REAL", "0")
        # ---- r2c
        cfg = pkg.make_config(dim, [M] + list(Ns) + [K], fp, pkg.FORWARD, pkg.R2C, inplace=inplace)
        plan = pkg.Plan(cfg, stream=_stream())
        names_f = plan.kernel_names
        if inplace:
            buf = torch.from_numpy(xin).cuda()
            plan.execute(buf)
            torch.cuda.synchronize()
            spec = buf.cpu().numpy().reshape(-1).view(cdtype(fp)).reshape(spec_shape)
            buf2 = torch.from_numpy(xin).cuda()
            plan.execute(buf2)
            assert torch.equal(buf, buf2)
        else:
            xd = torch.from_numpy(xin).cuda()
            yd = torch.zeros(spec_shape, dtype=ct, device="cuda")
            plan.execute(xd, yd)
            y1 = yd.clone()
            yd.zero_()
            plan.execute(xd, yd)
            assert torch.equal(y1, yd)
            spec = yd.cpu().numpy()
            if M == 1 and mode == "fused":
                big = torch.zeros(xin.size + 1, dtype=rt, device="cuda")
                big[1:] = xd.reshape(-1)
                yd.zero_()
                plan.execute(big[1:], yd)  # odd element offset
                assert rel_l2(yd.cpu().numpy(), ref) < TOL[fp]
        assert rel_l2(spec, ref) < TOL[fp], names_f
        plan.close()
        # ---- c2r
        cfg = pkg.make_config(dim, [M] + list(Ns) + [K], fp, pkg.BACKWARD, pkg.C2R, inplace=inplace)
        plan = pkg.Plan(cfg, stream=_stream())
        names_b = plan.kernel_names
        if inplace:
            raw = np.zeros(xin.nbytes, dtype=np.uint8)
            raw[: sp.nbytes] = sp.view(np.uint8).reshape(-1)
            buf = torch.from_numpy(raw).cuda()
            plan.execute(buf)
            torch.cuda.synchronize()
            back = buf.cpu().numpy().reshape(-1).view(rdtype(fp)).reshape(xin.shape)[..., :N1, :]
        else:
            sd = torch.from_numpy(sp).cuda()
            od = torch.zeros(x.shape, dtype=rt, device="cuda")
            plan.execute(sd, od)
            o1 = od.clone()
            od.zero_()
            plan.execute(sd, od)
            assert torch.equal(o1, od)
            assert torch.equal(sd, torch.from_numpy(sp).cuda()), "out-of-place c2r must not touch its input"
            back = od.cpu().numpy()
            if M == 1 and mode == "fused":
                big = torch.zeros(x.size + 1, dtype=rt, device="cuda")
                plan.execute(sd, big[1:])
                assert rel_l2(big[1:].cpu().numpy().reshape(x.shape), x.astype(np.float64) * np.prod(Ns)) < TOL[fp]
        assert rel_l2(back, x.astype(np.float64) * np.prod(Ns)) < TOL[fp], names_b
        plan.close()
        got[mode] = (spec.copy(), np.array(back, copy=True), names_f, names_b)
    assert any(n.startswith("bbfft_r2c2d") for n in got["fused"][2]), got["fused"][2]
    assert any(n.startswith("bbfft_c2r2d") for n in got["fused"][3]), got["fused"][3]
    if "unfused" in got:
        assert not any("2d" in n for n in got["unfused"][2] + got["unfused"][3])
        assert rel_l2(got["fused"][0], got["unfused"][0]) < TOL[fp]
        assert rel_l2(got["fused"][1], got["unfused"][1]) < TOL[fp]
