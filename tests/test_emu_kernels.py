"""CPU tier: the REAL kernel source (csrc/kernels/bbfft_kernels.cuh + the planner's stub) run
under the fiber emulator (tests/emu) and compared with the oracle.  Checks every index map --
stage addressing, digit reversal, staged copies, padding, edge guards -- without a GPU.  The
GPU tier repeats the comparison on hardware through the C ABI."""
import numpy as np
import pytest

import emu
from common import TOL, cdtype, random_complex, rdtype, rel_l2


def _run_c2c(pkg, oracle, M, N, K, fp, d, inplace=False, tune="", istride=None, ostride=None):
    rng = np.random.default_rng(M * 1000 + N * 7 + K)
    cfg = pkg.make_config(1, [M, N, K], fp, d, pkg.C2C, istride=istride, ostride=ostride, inplace=False)
    ocfg = oracle.make_config(1, [M, N, K], fp, d, 0, istride=istride, ostride=ostride, inplace=False)
    size_in = K * cfg.istride[2]
    size_out = K * cfg.ostride[2]
    x = random_complex(rng, (size_in,), fp)
    ref = np.zeros(size_out, dtype=x.dtype)
    oracle.dft(ocfg, x, ref)
    if inplace:
        y = x.copy()
        emu.run(cfg, y, None, tune)
    else:
        y = np.zeros(size_out, dtype=x.dtype)
        emu.run(cfg, x, y, tune)
    return rel_l2(y, ref)


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,N,K", [
    (16, 2, 5), (16, 7, 3), (16, 16, 3), (16, 30, 2), (16, 64, 3), (16, 105, 2), (16, 128, 2), (16, 343, 1),
    (16, 512, 1), (1, 64, 7), (1, 8, 40), (1, 15, 33), (1, 256, 3), (2, 12, 9), (3, 30, 5), (5, 64, 3),
    (17, 27, 2), (32, 25, 2), (64, 49, 1), (3, 363, 1), (1, 2, 1), (7, 1, 3),
])
def test_c2c_emulated_kernel_vs_oracle(pkg, oracle, fp, M, N, K):
    d = -1 if (M + N) % 2 else 1
    err = _run_c2c(pkg, oracle, M, N, K, fp, d, inplace=(N % 3 == 0))
    assert err < TOL[fp] * 0.1


@pytest.mark.parametrize("tune", ["R=4x16,T=4", "R=16x4,T=16", "R=2x4x8,T=8", "R=8x8,T=8,LD=1,ST=1", "R=8x8,T=4,BH=3",
                                  "R=64,T=1", "R=8x8,T=8,ML=4"])
def test_c2c_emulated_tuning_overrides(pkg, oracle, tune):
    """Every planner override produces a correct kernel (the auto-tuner explores these)."""
    assert _run_c2c(pkg, oracle, 16, 64, 5, 4, -1, tune=tune) < TOL[4] * 0.1


def test_c2c_emulated_nonpacked_strides(pkg, oracle):
    # reference test/c2c.cpp:68-82: strides (1, M+1, (M+1)(N+1)), K=33 (kept small here)
    M, N, K = 3, 16, 5
    s = [1, M + 1, (M + 1) * (N + 1)]
    assert _run_c2c(pkg, oracle, M, N, K, 4, -1, istride=s, ostride=s) < TOL[4] * 0.1
    assert _run_c2c(pkg, oracle, M, N, K, 8, 1, istride=s, ostride=[1, M, M * N]) < TOL[8] * 0.1


def _run_real(pkg, oracle, ttype, M, N, K, fp, inplace, pollute=False, tune=""):
    from common import addressed_mask, real_problem
    rng = np.random.default_rng(ttype * 7919 + M * 131 + N * 17 + K)
    d = -1 if ttype == 1 else 1
    ist, ost, x, odt, nout = real_problem(rng, pkg, ttype, M, N, K, fp, inplace, pollute=pollute)
    cfg = pkg.make_config(1, [M, N, K], fp, d, ttype, inplace=inplace)
    ocfg = oracle.make_config(1, [M, N, K], fp, d, ttype, inplace=inplace)
    if inplace:
        nbytes = max(x.nbytes, nout * np.dtype(odt).itemsize)
        raw = np.zeros(nbytes, np.uint8)
        raw[: x.nbytes] = x.view(np.uint8)
        ref = raw.copy()
        oracle.dft(ocfg, ref)
        emu.run(cfg, raw, None, tune)
        got, want = raw.view(odt)[:nout], ref.view(odt)[:nout]
    else:
        want = np.zeros(nout, odt)
        oracle.dft(ocfg, x, want)
        got = np.zeros(nout, odt)
        emu.run(cfg, x, got, tune)
    mask = addressed_mask(M, N // 2 + 1 if ttype == 1 else N, K, ost, nout)
    if not inplace:
        assert np.all(got[~mask] == 0), "kernel wrote outside the addressed elements"
    return rel_l2(got[mask], want[mask])


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("ttype", [1, 2])
@pytest.mark.parametrize("M,N,K", [(1, 256, 5), (1, 8, 40), (1, 30, 7), (16, 30, 3), (16, 64, 3), (4, 128, 2), (3, 10, 5),
                                   (1, 15, 7), (2, 15, 5), (16, 27, 3), (5, 11, 4), (1, 105, 3), (1, 2, 3), (16, 2, 3),
                                   (1, 512, 2), (16, 500, 1), (1, 3, 1)])
def test_real_emulated_kernel_vs_oracle(pkg, oracle, fp, ttype, M, N, K):
    """r2c / c2r, even N (half-length trick) and odd N (two-for-one, incl. odd K), in- and out-of-place."""
    for inplace in (False, True):
        d = pkg.describe(pkg.make_config(1, [M, N, K], fp, -1 if ttype == 1 else 1, ttype, inplace=inplace))
        if inplace and d["inplace_unsupported"]:
            continue
        assert _run_real(pkg, oracle, ttype, M, N, K, fp, inplace, pollute=(ttype == 2)) < TOL[fp] * 0.1


@pytest.mark.parametrize("ttype", [1, 2])
@pytest.mark.parametrize("M,N,K,tune", [(16, 64, 3, "RF=0"), (16, 64, 3, "RF=1"), (1, 256, 5, "RF=1,LD=0,ST=0"), (1, 256, 5, "RF=0,LD=0,ST=0"),
                                        (16, 27, 3, "RF=0"), (4, 105, 4, "RF=0,R=3x5x7"), (4, 105, 5, "RF=1,R=7x5x3,T=11"),
                                        (8, 500, 2, "RF=1,R=25x10,T=13"), (8, 500, 2, "RF=1,R=10x25,T=6"), (16, 48, 3, "RF=1,ML=8")])
def test_real_fused_and_separate_pass_emulated(pkg, oracle, ttype, M, N, K, tune):
    """The real pre/post pass fused into the first/last stage (RF=1, mirrored sub-FFT pairs) and run
    as a separate pass over shared memory (RF=0) are both planner choices; the tuner explores them."""
    for fp in (4, 8):
        for inplace in (False, True):
            d = pkg.describe(pkg.make_config(1, [M, N, K], fp, -1 if ttype == 1 else 1, ttype, inplace=inplace), tune)
            assert ("_rf1" in d["identifier"]) == ("RF=1" in tune)
            if inplace and d["inplace_unsupported"]:
                continue
            assert _run_real(pkg, oracle, ttype, M, N, K, fp, inplace, pollute=(ttype == 2), tune=tune) < TOL[fp] * 0.1


def test_real_inplace_lanes_cover_m(pkg):
    # strides that let the spectrum overlay the real rows make the lanes cover M (M <= 32), so the
    # cases the reference runs in place (test/r2c.cpp, golden c2r_f64_M16_N48_K3_ip) are supported
    for fp in (4, 8):
        for M in (2, 3, 16, 17, 32):
            for ttype, d in ((pkg.R2C, pkg.FORWARD), (pkg.C2R, pkg.BACKWARD)):
                for N in (48, 27, 441):
                    desc = pkg.describe(pkg.make_config(1, [M, N, 6], fp, d, ttype, inplace=True))
                    assert not desc["inplace_unsupported"], (fp, M, N, desc["identifier"])


def test_real_inplace_unsupported_flag(pkg):
    # a CTA must own every m of a k slice for real in-place transforms (reference
    # src/base/generator/small_batch_fft.cpp:38): with M beyond the lane count the plan refuses in-place
    d = pkg.describe(pkg.make_config(1, [33, 4, 2], 4, pkg.FORWARD, pkg.R2C, inplace=True))
    assert d["inplace_unsupported"]
    d = pkg.describe(pkg.make_config(1, [16, 4, 2], 4, pkg.FORWARD, pkg.R2C, inplace=True))
    assert not d["inplace_unsupported"]


@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,N", [(1, 8), (3, 12), (32, 8)])
def test_callbacks_emulated_bit_identical(pkg, fp, M, N):
    """Reference test/callback.cpp: a c2r plan with a zero-padding load callback and an r2c plan
    with a truncate+scale store callback equal the plain plans on the padded / full problem
    bit for bit (the FFT arithmetic is the same code; only the accessors differ)."""
    from callbacks import load_zero_pad_opencl, store_truncate_scale_opencl
    from common import cdtype, rdtype
    K = 4
    real = "float" if fp == 4 else "double"
    rng = np.random.default_rng(3)
    # ---- load callback (c2r of length 2N, upper half of the spectrum implied zero)
    N_ext = 2 * N
    ns_ref, ns = N_ext // 2 + 1, N // 2 + 1
    X = (rng.uniform(0, 1, (K, ns, M)) + 1j * rng.uniform(0, 1, (K, ns, M))).astype(cdtype(fp))
    X_ref = np.zeros((K, ns_ref, M), dtype=cdtype(fp))
    X_ref[:, :ns, :] = X
    strides = dict(istride=[1, M, M * ns_ref], ostride=[1, M, M * N_ext])
    cfg_ref = pkg.make_config(1, [M, N_ext, K], fp, pkg.BACKWARD, pkg.C2R, **strides)
    cfg = pkg.make_config(1, [M, N_ext, K], fp, pkg.BACKWARD, pkg.C2R,
                          callbacks=(load_zero_pad_opencl(real, M, ns_ref, ns), "load", None, "opencl"), **strides)
    x_ref = np.zeros(K * N_ext * M, dtype=rdtype(fp))
    x = np.zeros_like(x_ref)
    emu.run(cfg_ref, X_ref.reshape(-1), x_ref)
    _, d = emu.run(cfg, X.reshape(-1), x)
    assert d["identifier"].endswith("_load")
    assert np.array_equal(x, x_ref)
    # ---- store callback (r2c, keep N/4 rows scaled by 1/N)
    N2 = 4 * N
    ncut, nspec = N2 // 4, N2 // 2 + 1
    xin = rng.uniform(0, 1, K * N2 * M).astype(rdtype(fp))
    strides = dict(istride=[1, M, M * N2], ostride=[1, M, M * nspec])
    cfg_ref = pkg.make_config(1, [M, N2, K], fp, pkg.FORWARD, pkg.R2C, **strides)
    cfg = pkg.make_config(1, [M, N2, K], fp, pkg.FORWARD, pkg.R2C,
                          callbacks=(store_truncate_scale_opencl(real, M, nspec, ncut, 1.0 / N2), None, "store", "opencl"),
                          **strides)
    Y_ref = np.zeros(K * nspec * M, dtype=cdtype(fp))
    Y = np.zeros(K * ncut * M, dtype=cdtype(fp))
    emu.run(cfg_ref, xin, Y_ref)
    emu.run(cfg, xin, Y)
    want = (Y_ref.reshape(K, nspec, M)[:, :ncut, :] * rdtype(fp)(1.0 / N2)).astype(cdtype(fp))
    assert np.array_equal(Y.reshape(K, ncut, M), want)


# ---- fused 2d tile kernel (bbk::fft2d_tile) -------------------------------------------------
@pytest.mark.parametrize("fp", [4, 8])
@pytest.mark.parametrize("M,N1,N2,K,tune", [
    (1, 32, 32, 2, ""), (1, 64, 64, 1, ""), (1, 40, 30, 2, ""), (16, 8, 8, 2, ""), (3, 20, 18, 2, ""),
    (1, 128, 16, 1, ""), (1, 16, 128, 1, "RA=16,RB=2x8x8"), (2, 25, 49, 1, "TH=128"), (1, 36, 36, 1, "PADK=0"),
    (1, 33, 34, 2, ""),
])
def test_c2c_2d_emulated_tile_kernel_vs_oracle(pkg, oracle, fp, M, N1, N2, K, tune):
    """The fused 2d kernel: stage addressing in the tile, the sorted (digit-reversing) last stage
    of pass A, padding, ragged thread counts -- against the oracle's direct 2d DFT."""
    rng = np.random.default_rng(M * 1000 + N1 * 7 + N2 * 3 + K)
    d = -1 if (N1 + N2) % 3 else 1
    cfg = pkg.make_config(2, [M, N1, N2, K], fp, d, pkg.C2C, inplace=False)
    ocfg = oracle.make_config(2, [M, N1, N2, K], fp, d, 0, inplace=False)
    n = M * N1 * N2 * K
    x = random_complex(rng, (n,), fp)
    ref = np.zeros(n, dtype=x.dtype)
    oracle.dft(ocfg, x, ref)
    y = np.zeros(n, dtype=x.dtype)
    _, desc = emu.run(cfg, x, y, tune)
    assert desc["identifier"].startswith("bbfft_c2c2d")
    assert rel_l2(y, ref) < TOL[fp] * 0.1
    # in place: every global read of a tile precedes its first global write
    z = x.copy()
    emu.run(cfg, z, None, tune)
    assert np.array_equal(z, y)


@pytest.mark.parametrize("fp,M,N1,N2,K,tune,grid", [
    (4, 1, 32, 32, 5, "SG=24", 2), (8, 1, 32, 32, 5, "SG=-1", 2), (4, 1, 64, 64, 3, "SG=48", 1), (8, 1, 40, 30, 4, "SG=7", 3),
    (4, 16, 8, 8, 3, "SG=2", 2), (8, 3, 20, 18, 3, "SG=10", 1), (4, 1, 16, 128, 2, "RA=16,RB=2x8x8,SG=100", 1),
    (8, 1, 128, 16, 3, "SG=1,PS=1", 2), (4, 1, 32, 32, 4, "PS=1", 3), (8, 2, 25, 49, 2, "TH=128,SG=48", 1),
    (4, 1, 32, 32, 5, "SG=-1,BK=1", 2), (8, 1, 32, 32, 5, "SG=-1,BK=1", 2), (4, 1, 40, 30, 4, "SG=7,BK=1", 3),
])
def test_c2c_2d_emulated_staged_tile_kernel(pkg, oracle, fp, M, N1, N2, K, tune, grid):
    """The persistent tile kernels (asynchronous pipeline, PS=1; staging buffer for the leading rows of the
    next tile, SG=<rows>): a CTA walks several tiles, the first stage of pass A reads part of its input from
    the staging buffer -- bit-identical to the one-tile-per-CTA kernel, and right against the oracle."""
    rng = np.random.default_rng(M * 1000 + N1 * 7 + N2 * 3 + K + 11)
    d = -1 if (N1 + N2) % 3 else 1
    cfg = pkg.make_config(2, [M, N1, N2, K], fp, d, pkg.C2C, inplace=False)
    ocfg = oracle.make_config(2, [M, N1, N2, K], fp, d, 0, inplace=False)
    n = M * N1 * N2 * K
    x = random_complex(rng, (n,), fp)
    ref = np.zeros(n, dtype=x.dtype)
    oracle.dft(ocfg, x, ref)
    plain = np.zeros(n, dtype=x.dtype)
    base_tune = ",".join(t for t in tune.split(",") if not t.startswith(("SG=", "PS=", "BK=")))
    emu.run(cfg, x, plain, base_tune)
    y = np.zeros(n, dtype=x.dtype)
    _, desc = emu.run(cfg, x, y, tune, grid=grid)
    assert "_sg" in desc["identifier"] or "_ps" in desc["identifier"]
    assert rel_l2(y, ref) < TOL[fp] * 0.1
    assert np.array_equal(y, plain)
    z = x.copy()
    emu.run(cfg, z, None, tune, grid=grid)
    assert np.array_equal(z, y)


@pytest.mark.parametrize("fp,inplace,M,N1,N2,K", [
    (4, False, 1, 64, 32, 2), (8, True, 1, 64, 32, 2), (4, True, 1, 32, 64, 1), (8, False, 3, 20, 36, 2), (4, True, 3, 20, 36, 2),
    (4, False, 2, 12, 160, 1), (8, True, 16, 8, 16, 2), (4, False, 16, 8, 16, 2), (4, False, 1, 128, 16, 1), (8, True, 1, 128, 16, 1),
    (8, False, 1, 40, 60, 2), (4, True, 1, 40, 60, 2), (4, False, 1, 6, 400, 1), (8, True, 1, 4, 512, 1), (4, False, 1, 4, 512, 1)])
def test_real_2d_emulated_fused_tile_kernel(pkg, fp, inplace, M, N1, N2, K):
    """The fused real tile kernels (bbk::fft2d_tile_real_cta): r2c = half-length pass along n1 from the real rows,
    split on the pairs (i, N1/2 - i), column pass; c2r = the mirror image.  Against numpy's rfft2 / irfft2 in
    float64, out of place and in the padded in-place layout."""
    rng = np.random.default_rng(M * 100 + N1 * 7 + N2 + K)
    ns = N1 // 2 + 1
    n1r = 2 * ns if inplace else N1
    x = rng.uniform(-1, 1, (K, N2, N1, M)).astype(rdtype(fp))
    ref = np.fft.fft(np.fft.rfft(x.astype(np.float64), axis=2), axis=1)
    # ---- r2c
    cfg = pkg.make_config(2, [M, N1, N2, K], fp, pkg.FORWARD, pkg.R2C, inplace=inplace)
    xin = np.zeros((K, N2, n1r, M), dtype=rdtype(fp))
    xin[:, :, :N1, :] = x
    if inplace:
        buf = xin.reshape(-1).copy()
        _, desc = emu.run(cfg, buf, None)
        spec = buf.view(cdtype(fp)).reshape(K, N2, ns, M)
    else:
        spec = np.zeros((K, N2, ns, M), dtype=cdtype(fp))
        _, desc = emu.run(cfg, xin.reshape(-1), spec.reshape(-1))
    assert desc["identifier"].startswith("bbfft_r2c2d")
    assert rel_l2(spec, ref) < TOL[fp] * 0.1
    # ---- c2r from the exact spectrum, imaginary part of X[0] polluted (reference test/r2c.cpp:310-324)
    cfg = pkg.make_config(2, [M, N1, N2, K], fp, pkg.BACKWARD, pkg.C2R, inplace=inplace)
    sp = ref.astype(cdtype(fp))
    sp[:, 0, 0, :] += 0.37j  # ignored
    want = x.astype(np.float64) * (N1 * N2)
    if inplace:
        buf = np.zeros(K * N2 * n1r * M, dtype=rdtype(fp))
        buf.view(cdtype(fp))[:] = sp.reshape(-1)
        _, desc = emu.run(cfg, buf, None)
        back = buf.reshape(K, N2, n1r, M)[:, :, :N1, :]
    else:
        back = np.zeros((K, N2, N1, M), dtype=rdtype(fp))
        _, desc = emu.run(cfg, sp.reshape(-1).copy(), back.reshape(-1))
    assert desc["identifier"].startswith("bbfft_c2r2d")
    assert rel_l2(back, want) < TOL[fp] * 0.1


@pytest.mark.parametrize("mutation", ["no barrier before the staging copies", "no wait for the tile copies"])
def test_emulator_catches_async_copy_hazards(pkg, monkeypatch, tmp_path, mutation):
    """The emulator models cp.async / cp.async.bulk pessimistically (destination poisoned from issue to wait, the
    issue counts as a write of the issuing thread).  Two mutations of the staged tile kernel prove that the model
    bites: without the barrier in front of the staging copies the race checker reports the threads that still
    read the buffer; without the wait the first stage consumes poison and the result is NaN."""
    import os
    src = open(os.path.join(emu._KERNELS, "bbfft_kernels.cuh")).read()
    if mutation.startswith("no barrier"):
        old = "        BBK_SYNC(); // the tile (or the staging buffer) has left shared memory\n        after_loads();"
        new = "        after_loads();"
    else:
        old = "        async_wait_all();\n        if (bulk) {\n            mbar_wait(sm, C::STG_OFF + C::STG, phase);"
        new = "        if (bulk) {\n            mbar_wait(sm, C::STG_OFF + C::STG, phase);"
    assert src.count(old) == 1
    (tmp_path / "bbfft_kernels.cuh").write_text(src.replace(old, new))
    monkeypatch.setenv("BBFFT_EMU_KERNELS_DIR", str(tmp_path))
    monkeypatch.setattr(emu, "RACECHECK", True)
    cfg = pkg.make_config(2, [1, 32, 32, 5], 4, -1, pkg.C2C, inplace=False)
    rng = np.random.default_rng(5)
    x = random_complex(rng, (32 * 32 * 5,), 4)
    y = np.zeros_like(x)
    if mutation.startswith("no barrier"):
        with pytest.raises(RuntimeError, match="shared-memory race"):
            emu.run(cfg, x, y, "SG=24", grid=2)
    else:
        emu.run(cfg, x, y, "SG=24", grid=2)
        assert np.isnan(y.view(np.float32)).any()
    # the unmutated header passes the same run
    monkeypatch.delenv("BBFFT_EMU_KERNELS_DIR")
    y2 = np.zeros_like(x)
    emu.run(cfg, x, y2, "SG=24", grid=2)
    assert not np.isnan(y2.view(np.float32)).any()


# ---- chained nd kernel (bbk::chain): all steps of a 2d/3d plan in one persistent launch ----------
@pytest.mark.parametrize("desc,fp,shape_np,kind,kblock,epochs", [
    ("dcfo32x32x32*3", 8, (3, 32, 32, 32, 1), "c2c", 2, 1),      # fused tile step + 1d pass
    ("dcfo32x32x32*5", 8, (5, 32, 32, 32, 1), "c2c", 2, 2),      # second launch: counters keep counting
    ("srfo128x64*6", 4, (6, 64, 128, 1), "r2c", 4, 1),            # r2c pass + c2c pass with M' = 65
    ("scfo1024x64*3", 4, (3, 64, 1024, 1), "c2c", 1, 1),          # two 1d passes, M = 1 first
])
def test_chain_kernel_emulated(pkg, desc, fp, shape_np, kind, kblock, epochs):
    """One emulated CTA walks every work item of the chain in its dependency order (a wait that
    would spin on the GPU fails the run): index maps, slab/step decoding, counters."""
    cfg = pkg.parse_descriptor(desc)
    rng = np.random.default_rng(1)
    cdt = np.complex64 if fp == 4 else np.complex128
    axes = tuple(range(1, len(shape_np) - 1))
    if kind == "r2c":
        x = rng.standard_normal(shape_np).astype(np.float32 if fp == 4 else np.float64)
        ref = np.fft.rfftn(x.astype(np.float64), axes=axes)
    else:
        x = (rng.standard_normal(shape_np) + 1j * rng.standard_normal(shape_np)).astype(cdt)
        ref = np.fft.fftn(x.astype(np.complex128), axes=axes)
    y = np.zeros(ref.shape, cdt)
    _, d, done = emu.run_chain(cfg, x.reshape(-1), y.reshape(-1), kblock=kblock, epochs=epochs)
    assert d["identifier"].startswith("bbfft_chain")
    assert rel_l2(y, ref) < TOL[fp] * 0.1
    K = shape_np[0]
    # every step but the last signals once per CTA and launch
    for s in range(d["n_steps"] - 1):
        assert np.all(done[s * K:(s + 1) * K] == epochs * d["per_k"][s])


def test_chain_planning_falls_back(pkg):
    # steps that cannot share one CTA shape are launched one by one: no chain description
    for desc in ("scfo3.6x5x4*2", "scfo128x128*64"):
        with pytest.raises(pkg.BadConfiguration):
            pkg.describe_chain(pkg.parse_descriptor(desc))
    d = pkg.describe_chain(pkg.parse_descriptor("dcfo64x64x64*64"))
    assert d["step_tile"] == [1, 0] and d["per_k"] == [64, 128] and d["threads"] == 256


# ---- large prime factors: cooperative direct-DFT stages (bbk::run_stage_direct) ---------------------
@pytest.mark.parametrize("M,N,K,fp", [(16, 37, 3, 4), (16, 127, 2, 4), (3, 101, 4, 8), (1, 67, 5, 8), (16, 254, 2, 4),
                                      (2, 424, 2, 4), (1, 509, 2, 4), (4, 509, 1, 8)])
def test_c2c_large_prime_factor_emulated(pkg, oracle, M, N, K, fp):
    """Every N works, as in the reference (which unrolls a prime length into one work-item): a prime
    factor beyond the in-register butterflies becomes a direct-DFT stage out of shared memory."""
    d = pkg.describe(pkg.make_config(1, [M, N, K], fp, -1, 0, inplace=False))
    assert max(d["radix"]) > 31 and d["smem_bytes"] > 0
    for direction in (-1, 1):
        assert _run_c2c(pkg, oracle, M, N, K, fp, direction) < TOL[fp] * 0.2


@pytest.mark.parametrize("ttype", [1, 2])
@pytest.mark.parametrize("M,N,K,fp", [(16, 74, 3, 4), (16, 202, 2, 4), (1, 254, 4, 4), (3, 127, 4, 8), (1, 127, 5, 4), (2, 106, 3, 8)])
def test_real_large_prime_factor_emulated(pkg, oracle, ttype, M, N, K, fp):
    for inplace in (False, True):
        d = pkg.describe(pkg.make_config(1, [M, N, K], fp, -1 if ttype == 1 else 1, ttype, inplace=inplace))
        assert "_rf0" in d["identifier"]  # a direct stage is never the fused (mirrored, in-register) one
        if inplace and d["inplace_unsupported"]:
            continue
        assert _run_real(pkg, oracle, ttype, M, N, K, fp, inplace, pollute=(ttype == 2)) < TOL[fp] * 0.2


# ---- robustness of the kernels against what the hardware does not promise ----------------------------
@pytest.mark.parametrize("desc,tune", [
    ("scfo16.64*5", ""), ("scfo1.64*40", ""), ("dcfo3.105*7", ""), ("scfo16.254*3", ""), ("scfo1.101*9", ""),
    ("srfo16.256*3", ""), ("srbo16.256*3", ""), ("srfo1.256*5", ""), ("srbo1.256*5", ""), ("srbi1.16*4", ""), ("srbo1.16*4", ""),
    ("srbo32.424*4", ""), ("srbo1.424*4", ""), ("srfo3.27*5", ""), ("drbo16.127*4", ""), ("srfo16.64*3", "RF=0"),
    ("dcfo32x32*3", ""), ("scfo16.32x48*2", ""), ("srfi16.30*5", ""), ("srbi16.30*5", ""), ("drfi3.27*5", ""), ("srfi32.100*3", ""),
    ("srbi1.256*6", ""), ("drbo5.11*4", ""), ("scfo17.12*3", ""),
])
def test_output_independent_of_smem_garbage_and_thread_order(pkg, monkeypatch, desc, tune):
    """Shared memory is not zeroed between CTAs and warps run in any order: the result must not
    change when the emulator pre-fills shared memory with another byte or runs the CTAs, and the threads
    of every barrier interval, last-to-first (complements BBFFT_EMU_RACECHECK=1)."""
    cfg = pkg.parse_descriptor(desc)
    n = 1
    for d in range(cfg.dim + 2):
        n *= cfg.shape[d]
    rng = np.random.default_rng(11)
    x = rng.uniform(0, 1, 2 * n + 64).astype(np.float32 if cfg.fp == 4 else np.float64)
    outs = []
    for env in ({}, {"BBFFT_EMU_SMEM_FILL": "0"}, {"BBFFT_EMU_SMEM_FILL": "255", "BBFFT_EMU_ORDER": "reverse"}):
        for k in ("BBFFT_EMU_SMEM_FILL", "BBFFT_EMU_ORDER"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        y = np.zeros(2 * n + 64, dtype=x.dtype)
        if desc[3] == "i":
            y[:] = x  # in-place configurations transform their own buffer
            emu.run(cfg, y, None, tune)
        else:
            emu.run(cfg, x.copy(), y, tune)
        outs.append(y)
    assert np.array_equal(outs[0], outs[1], equal_nan=True)
    assert np.array_equal(outs[0], outs[2], equal_nan=True)


@pytest.mark.parametrize("M,N,K,fp", [(16, 2048, 1, 4), (16, 2048, 1, 8), (1, 4096, 2, 4)])
def test_c2c_beyond_512_emulated(pkg, oracle, M, N, K, fp):
    """N > 512 stays one kernel while a transform fits shared memory (the reference's f2fft kernel has
    the same bound); with many batch lanes the planner narrows the lanes until the CTA's rows fit."""
    d = pkg.describe(pkg.make_config(1, [M, N, K], fp, -1, 0, inplace=False))
    assert d["smem_bytes"] <= 227 * 1024 and (M == 1 or d["batch_lanes"] < 128 // (2 * fp) or N * 128 <= 227 * 1024)
    assert _run_c2c(pkg, oracle, M, N, K, fp, -1) < TOL[fp] * 0.2
