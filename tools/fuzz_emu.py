"""Randomised parity check of planned kernels under the CPU emulator against the long-double oracle:
random (type, precision, direction, M, N, K, strides, in/out-of-place) -- the planner picks the
kernel -- optionally with random planner overrides.  Test infrastructure, no GPU needed.
Usage: python tools/fuzz_emu.py --n 200 --seed 1 [--racecheck]"""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--racecheck", action="store_true")
    ap.add_argument("--maxn", type=int, default=160)
    ap.add_argument("--nd", type=int, default=0, help="additional random 2d / 3d c2c configurations (tile kernel, chain)")
    ap.add_argument("--real2d", type=int, default=0, help="additional random 2d r2c / c2r configurations (fused real tile kernels)")
    args = ap.parse_args()
    if args.racecheck:
        os.environ["BBFFT_EMU_RACECHECK"] = "1"
    import numpy as np
    import emu
    from oracle import oracle
    from common import TOL, rel_l2
    pkg = emu.pkg
    rnd = random.Random(args.seed)
    bad = 0
    for it in range(args.n):
        ttype = rnd.choice([0, 0, 1, 2])
        fp = rnd.choice([4, 8])
        M = rnd.choice([1, 1, 2, 3, 4, 5, 7, 8, 12, 16, 17, 24, 32, 33])
        N = rnd.randint(2, args.maxn)
        K = rnd.randint(1, 9)
        d = rnd.choice([-1, 1]) if ttype == 0 else (-1 if ttype == 1 else 1)
        inplace = rnd.random() < 0.4
        nspec = N // 2 + 1
        # default strides, or padded ones (c2c only: the real layouts are pinned by the in-place rule)
        kw = {}
        if ttype == 0 and rnd.random() < 0.3:
            s1 = M + rnd.randint(0, 3)
            s2 = s1 * N + rnd.randint(0, 5)
            kw = dict(istride=[1, s1, s2], ostride=[1, s1, s2] if inplace else [1, M, M * N])
        try:
            cfg = pkg.make_config(1, [M, N, K], fp, d, ttype, inplace=inplace, **kw) if not kw else \
                pkg.make_config(1, [M, N, K], fp, d, ttype, **kw)
            desc = pkg.describe(cfg)
        except Exception as ex:
            print("plan rejected:", ttype, fp, M, N, K, inplace, kw, str(ex)[:80])
            bad += 1
            continue
        if inplace and desc["inplace_unsupported"]:
            continue
        ocfg = oracle.make_config(1, [M, N, K], fp, d, ttype, inplace=inplace, **kw) if not kw else \
            oracle.make_config(1, [M, N, K], fp, d, ttype, **kw)
        rng = np.random.default_rng(it)
        rdt = np.float32 if fp == 4 else np.float64
        cdt = np.complex64 if fp == 4 else np.complex128
        ist, ost = list(cfg.istride)[:3], list(cfg.ostride)[:3]
        in_elems = ist[2] * K + 8
        out_elems = ost[2] * K + 8
        if ttype == 1:
            x = rng.standard_normal(max(in_elems, 2 * out_elems)).astype(rdt)
        else:
            x = (rng.standard_normal(max(in_elems, out_elems)) + 1j * rng.standard_normal(max(in_elems, out_elems))).astype(cdt)
        if ttype == 2 and N % 2 == 0:
            # a valid c2r input has a real Nyquist bin (imag(X[0]) may be polluted: it is ignored by
            # definition, reference test/r2c.cpp:310-324; imag(X[N/2]) is not)
            for k in range(K):
                for m in range(M):
                    idx = m + (N // 2) * ist[1] + k * ist[2]
                    x[idx] = x[idx].real
        try:
            if inplace:
                buf, ref = x.copy(), x.copy()
                emu.run(cfg, buf, None)
                oracle.dft(ocfg, ref)
                got, want = buf.view(rdt), ref.view(rdt)
            else:
                odt = rdt if ttype == 2 else cdt
                got = np.zeros(out_elems, odt)
                want = np.zeros(out_elems, odt)
                emu.run(cfg, x, got)
                oracle.dft(ocfg, x, want)
                got, want = got.view(rdt), want.view(rdt)
            # in-place buffers keep untouched input in their gaps: compare where the oracle wrote or left equal
            err = rel_l2(got, want)
        except Exception as ex:
            print("FAILED to run:", ttype, fp, M, N, K, inplace, kw, desc["identifier"], str(ex)[:120])
            bad += 1
            continue
        ok = err < TOL[fp] * 0.5
        if not ok:
            bad += 1
        print("%s type=%d fp=%d M=%d N=%d K=%d dir=%d inplace=%d %s err=%.2e %s" % (
            "ok  " if ok else "BAD ", ttype, fp, M, N, K, d, inplace, kw or "", err, desc["identifier"][6:70]), flush=True)
    # ---- 2d / 3d: fused tile kernel and chained plans (c2c, default layout)
    for it in range(args.nd):
        fp = rnd.choice([4, 8])
        dim = rnd.choice([2, 2, 3])
        M = rnd.choice([1, 1, 2, 3])
        Ns = [rnd.choice([4, 6, 8, 9, 10, 12, 15, 16, 20, 24, 27, 32, 36, 48, 64]) for _ in range(dim)]
        K = rnd.randint(1, 4)
        d = rnd.choice([-1, 1])
        cfg = pkg.make_config(dim, [M] + Ns + [K], fp, d, 0, inplace=False)
        cdt = np.complex64 if fp == 4 else np.complex128
        rng = np.random.default_rng(1000 + it)
        shape_np = (K,) + tuple(reversed(Ns)) + (M,)
        x = (rng.standard_normal(shape_np) + 1j * rng.standard_normal(shape_np)).astype(cdt)
        axes = tuple(range(1, dim + 1))
        ref = np.fft.fftn(x.astype(np.complex128), axes=axes) if d < 0 else np.fft.ifftn(x.astype(np.complex128), axes=axes) * np.prod(Ns)
        y = np.zeros_like(x)
        what = None
        try:
            if dim == 2:
                try:
                    _, desc = emu.run(cfg, x.reshape(-1), y.reshape(-1))
                    what = desc["identifier"]
                except pkg.BadConfiguration:
                    pass
            if what is None:
                try:
                    _, desc, _ = emu.run_chain(cfg, x.reshape(-1), y.reshape(-1), kblock=rnd.randint(1, 3), epochs=rnd.randint(1, 2))
                    what = desc["identifier"]
                except pkg.BadConfiguration:
                    continue  # neither a single fused kernel nor chainable: one launch per step, covered by the 1d runs
        except Exception as ex:
            print("FAILED to run nd:", fp, M, Ns, K, str(ex)[:160])
            bad += 1
            continue
        err = rel_l2(y, ref)
        ok = err < TOL[fp] * 0.5
        bad += 0 if ok else 1
        print("%s nd fp=%d M=%d N=%s K=%d dir=%d err=%.2e %s" % ("ok  " if ok else "BAD ", fp, M, Ns, K, d, err, what[:70]), flush=True)
    # ---- real 2d: fused r2c / c2r tile kernels (even N1, default layouts, in and out of place)
    for it in range(args.real2d):
        fp = rnd.choice([4, 8])
        M = rnd.choice([1, 1, 2, 3, 4])
        N1 = rnd.choice([4, 6, 8, 10, 12, 16, 18, 20, 24, 30, 32, 36, 40, 48, 50, 64, 96, 128])
        N2 = rnd.choice([8, 9, 12, 15, 16, 20, 24, 27, 32, 36, 48, 64, 100, 128])
        K = rnd.randint(1, 3)
        inplace = rnd.choice([False, True])
        fwd = rnd.choice([True, False])
        ns = N1 // 2 + 1
        n1r = 2 * ns if inplace else N1
        rdt, cdt = (np.float32, np.complex64) if fp == 4 else (np.float64, np.complex128)
        rng = np.random.default_rng(5000 + it)
        x = rng.uniform(-1, 1, (K, N2, N1, M)).astype(rdt)
        spec = np.fft.fft(np.fft.rfft(x.astype(np.float64), axis=2), axis=1)
        cfg = pkg.make_config(2, [M, N1, N2, K], fp, -1 if fwd else 1, 1 if fwd else 2, inplace=inplace)
        try:
            if fwd:
                xin = np.zeros((K, N2, n1r, M), dtype=rdt)
                xin[:, :, :N1, :] = x
                if inplace:
                    buf = xin.reshape(-1).copy()
                    _, desc = emu.run(cfg, buf, None)
                    got = buf.view(cdt).reshape(K, N2, ns, M)
                else:
                    got = np.zeros((K, N2, ns, M), dtype=cdt)
                    _, desc = emu.run(cfg, xin.reshape(-1), got.reshape(-1))
                want = spec
            else:
                sp = spec.astype(cdt)
                if inplace:
                    buf = np.zeros(K * N2 * n1r * M, dtype=rdt)
                    buf.view(cdt)[:] = sp.reshape(-1)
                    _, desc = emu.run(cfg, buf, None)
                    got = buf.reshape(K, N2, n1r, M)[:, :, :N1, :]
                else:
                    got = np.zeros((K, N2, N1, M), dtype=rdt)
                    _, desc = emu.run(cfg, sp.reshape(-1).copy(), got.reshape(-1))
                want = x.astype(np.float64) * (N1 * N2)
        except pkg.BadConfiguration:
            continue  # not a single fused kernel (tile too small or too large): one launch per mode, covered by the 1d runs
        except Exception as ex:
            print("FAILED to run real 2d:", fp, M, N1, N2, K, str(ex)[:160])
            bad += 1
            continue
        err = rel_l2(got, want)
        ok = err < TOL[fp] * 0.5
        bad += 0 if ok else 1
        print("%s real2d %s fp=%d M=%d N=%dx%d K=%d inplace=%d err=%.2e %s" % (
            "ok  " if ok else "BAD ", "r2c" if fwd else "c2r", fp, M, N1, N2, K, inplace, err, desc["identifier"][:70]), flush=True)
    print("done: %d problems" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
