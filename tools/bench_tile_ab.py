"""A/B of the fused tile kernel's variants on BASELINE config 4 (needs a GPU).

Every variant is planned and timed in one process, interleaved (variant order rotates between rounds), with an
L2 flush between timed launches for the L2-resident shapes.  2d plans take a tune string; the 3d plan takes the
environment switch (its tile step is planned inside nd_plan).
Usage: python tools/bench_tile_ab.py [--rounds 7]
"""
import argparse
import importlib
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("double-batched-fft-library_b200")
PEAK = 6534.5
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def make(dims, K, fp, tune="", env=None, M=1):
    cfg = pkg.make_config(len(dims), [M] + list(dims) + [K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
    saved = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        return pkg.Plan(cfg, stream=torch.cuda.current_stream().cuda_stream, tune=tune)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


WARM = False


def warm(dims, K, fp, variants, M=1):
    """CPU-only box: compile every variant's tile kernel into the kernel cache (3d: the tile step is the 2d
    kernel over (n1, n2) with K * n3 tiles)."""
    d2 = list(dims[:2])
    k2 = K * (dims[2] if len(dims) == 3 else 1)
    cfg = pkg.make_config(2, [1] + d2 + [k2], fp, pkg.FORWARD, pkg.C2C, inplace=False)
    if len(dims) == 1:
        cfg = pkg.make_config(1, [M, dims[0], K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
    for name, tune, env in variants:
        saved = {k: os.environ.get(k) for k in (env or {})}
        os.environ.update(env or {})
        try:
            d = pkg.describe(cfg, tune)
            pkg.compile_to_cubin(d["source"])
            print("  warmed %-28s %s" % (name, d["identifier"]))
        except Exception as e:
            print("  %-28s refused: %s" % (name, str(e)[:120]))
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


def ab(title, dims, K, fp, variants, rounds, M=1):
    if WARM:
        return warm(dims, K, fp, variants, M)
    n = M
    for d in dims:
        n *= d
    rdt = torch.float32 if fp == 4 else torch.float64
    x = torch.view_as_complex(torch.rand(K * n, 2, dtype=rdt, device="cuda"))
    y = torch.empty_like(x)
    nbytes = 2.0 * 2 * fp * n * K
    plans, ref = [], None
    for name, tune, env in variants:
        try:
            p = make(dims, K, fp, tune, env, M)
        except Exception as e:  # a variant the planner refuses is reported, not fatal
            print("  %-28s refused: %s" % (name, str(e)[:120]))
            continue
        p.execute(x, y)
        torch.cuda.synchronize()
        if ref is None:
            ref = y.clone()
            same = True
        else:
            same = bool(torch.equal(ref, y))
        plans.append((name, p, same))
    times = {name: [] for name, _, _ in plans}
    for r in range(rounds):
        order = plans[r % len(plans):] + plans[:r % len(plans)]
        for name, p, _ in order:
            for _ in range(2):
                p.execute(x, y)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                p.execute(x, y)
            e1.record()
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1) / 4 * 1e3)
    print("%s  (%.0f MiB in)" % (title, nbytes / 2 / 2**20))
    for name, p, same in plans:
        t = statistics.median(times[name])
        print("  %-28s %9.2f us  %7.1f GB/s  %.3f of peak  bit-identical %s  %s" % (
            name, t, nbytes / t / 1e3, nbytes / t / 1e3 / PEAK, same, p.kernel_names[0][:90]))
        p.close()
    sys.stdout.flush()


def ab_real(title, dims, K, fp, ttype, variants, rounds):
    """r2c / c2r nd plans, out of place: `variants` = (name, tune, env); algorithmic bytes = real + spectrum tensor."""
    n = 1
    for d in dims:
        n *= d
    nspec = (dims[0] // 2 + 1) * (n // dims[0])
    rdt = torch.float32 if fp == 4 else torch.float64
    cdt = torch.complex64 if fp == 4 else torch.complex128
    fwd = ttype == "r2c"
    cfg = pkg.make_config(len(dims), [1] + list(dims) + [K], fp, pkg.FORWARD if fwd else pkg.BACKWARD,
                          pkg.R2C if fwd else pkg.C2R, inplace=False)
    if WARM:
        for name, tune, env in variants:
            if len(dims) != 2:
                continue
            saved = {k: os.environ.get(k) for k in (env or {})}
            os.environ.update(env or {})
            try:
                d = pkg.describe(cfg, tune)
                pkg.compile_to_cubin(d["source"])
                print("  warmed %-28s %s" % (name, d["identifier"]))
            except Exception as e:
                print("  %-28s not a single kernel: %s" % (name, str(e)[:80]))
            finally:
                for k, v in saved.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
        return
    xr = torch.rand(K * n, dtype=rdt, device="cuda")
    xc = torch.view_as_complex(torch.rand(K * nspec, 2, dtype=rdt, device="cuda"))
    src, dst = (xr, torch.empty_like(xc)) if fwd else (xc, torch.empty_like(xr))
    nbytes = float(xr.numel() * fp + xc.numel() * 2 * fp)
    plans = []
    for name, tune, env in variants:
        saved = {k: os.environ.get(k) for k in (env or {})}
        os.environ.update(env or {})
        try:
            p = pkg.Plan(cfg, stream=torch.cuda.current_stream().cuda_stream, tune=tune)
        except Exception as e:
            print("  %-28s refused: %s" % (name, str(e)[:120]))
            continue
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        p.execute(src, dst)
        torch.cuda.synchronize()
        plans.append((name, p))
    times = {name: [] for name, _ in plans}
    for r in range(rounds):
        order = plans[r % len(plans):] + plans[:r % len(plans)]
        for name, p in order:
            for _ in range(2):
                p.execute(src, dst)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                p.execute(src, dst)
            e1.record()
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1) / 4 * 1e3)
    print("%s  (%.0f MiB real + spectrum)" % (title, nbytes / 2**20))
    for name, p in plans:
        t = statistics.median(times[name])
        print("  %-28s %9.2f us  %7.1f GB/s  %.3f of peak  %d launch(es)  %s" % (
            name, t, nbytes / t / 1e3, nbytes / t / 1e3 / PEAK, p.launches_per_execute, p.kernel_names[0][:80]))
        p.close()
    sys.stdout.flush()


REAL_SWEEP = (2, 3, 4, 7, 8, 15, 16, 27, 32, 49, 64, 100, 105, 128, 135, 200, 243, 256, 315, 343, 384, 400, 441, 480, 500, 512)


def x2_sweep(rounds):
    """Packed fp32 adds per kernel: every fp32 size of the c2c sweep and of the r2c / c2r M=16 sweep, planned with and
    without X2=1, timed interleaved; prints the ratio (> 1: the packed build is faster)."""
    aot = importlib.import_module("double-batched-fft-library_b200.aot")
    jobs = [("c2c", n) for n in aot.smooth_sizes()] + [(t, n) for t in ("r2c", "c2r") for n in REAL_SWEEP if n > 2]
    if os.environ.get("SHARD"):  # "i/n": every n-th job (parallel warm-up on the build box)
        i, n_sh = map(int, os.environ["SHARD"].split("/"))
        jobs = jobs[i::n_sh]
    for ttype, n in jobs:
        if ttype == "c2c":
            K = (1 << 30) // (16 * n * 8)
            cfg = pkg.make_config(1, [16, n, K], 4, pkg.FORWARD, pkg.C2C, inplace=False)
        else:
            K = (1 << 29) // (16 * n * 4)
            fwd = ttype == "r2c"
            cfg = pkg.make_config(1, [16, n, K], 4, pkg.FORWARD if fwd else pkg.BACKWARD, pkg.R2C if fwd else pkg.C2R, inplace=False)
        if WARM:
            for tune in ("X2=0", "X2=1"):
                try:
                    pkg.compile_to_cubin(pkg.describe(cfg, tune)["source"])
                except Exception as e:
                    print("warm failed", ttype, n, tune, str(e)[:80])
            continue
        if ttype == "c2c":
            src = torch.view_as_complex(torch.rand(K * n * 16, 2, dtype=torch.float32, device="cuda"))
            dst = torch.empty_like(src)
        else:
            xr = torch.rand(K * n * 16, dtype=torch.float32, device="cuda")
            xc = torch.view_as_complex(torch.rand(K * (n // 2 + 1) * 16, 2, dtype=torch.float32, device="cuda"))
            src, dst = (xr, torch.empty_like(xc)) if ttype == "r2c" else (xc, torch.empty_like(xr))
        plans = [pkg.Plan(cfg, stream=torch.cuda.current_stream().cuda_stream, tune=t) for t in ("X2=0", "X2=1")]
        outs = []
        for p in plans:
            p.execute(src, dst)
            torch.cuda.synchronize()
            outs.append(dst.clone())
        same = bool(torch.equal(outs[0], outs[1]))
        times = [[], []]
        for r in range(rounds):
            for i in ((0, 1) if r % 2 == 0 else (1, 0)):
                plans[i].execute(src, dst)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    plans[i].execute(src, dst)
                e1.record()
                torch.cuda.synchronize()
                times[i].append(e0.elapsed_time(e1) / 3 * 1e3)
        t0, t1 = statistics.median(times[0]), statistics.median(times[1])
        print("%s f32 N=%-4d plain %8.2f us  x2 %8.2f us  ratio %.3f  bit-identical %s" % (ttype, n, t0, t1, t0 / t1, same))
        sys.stdout.flush()
        for p in plans:
            p.close()
        del src, dst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=7)
    ap.add_argument("--warm", action="store_true", help="no GPU: compile the variants into BBFFT_CUDA_KERNEL_CACHE")
    ap.add_argument("--which", default="tile", help="tile | x2")
    a = ap.parse_args()
    global WARM
    WARM = a.warm
    if a.which == "x2sweep":
        return x2_sweep(a.rounds)
    if a.which == "real":
        unf = {"BBFFT_CUDA_ND_FUSE_REAL": "0"}
        for tt in ("r2c", "c2r"):
            ab_real("2d %s f32 128x128 K=8192" % tt, (128, 128), 8192, 4, tt, [
                ("one launch per mode", "", unf), ("fused", "", None), ("fused MB=2", "MB=2", None),
                ("fused TH=256,MB=2", "TH=256,MB=2", None), ("fused TH=256,MB=3", "TH=256,MB=3", None),
                ("fused TH=1024", "TH=1024", None)], a.rounds)
            ab_real("2d %s f64 64x64 K=16384" % tt, (64, 64), 16384, 8, tt, [
                ("one launch per mode", "", unf), ("fused", "", None), ("fused TH=256", "TH=256", None),
                ("fused TH=256,MB=3", "TH=256,MB=3", None), ("fused TH=128,MB=5", "TH=128,MB=5", None)], a.rounds)
            ab_real("2d %s f32 64x64 K=32768" % tt, (64, 64), 32768, 4, tt, [
                ("one launch per mode", "", unf), ("fused", "", None), ("fused TH=256", "TH=256", None), ("fused TH=64", "TH=64", None)], a.rounds)
            ab_real("2d %s f32 256x64 K=8192" % tt, (256, 64), 8192, 4, tt, [("one launch per mode", "", unf), ("fused", "", None), ("fused MB=2", "MB=2", None)], a.rounds)
            ab_real("2d %s f64 128x128 K=4096" % tt, (128, 128), 4096, 8, tt, [("one launch per mode", "", unf), ("fused", "", None)], a.rounds)
            ab_real("3d %s f64 64^3 K=64" % tt, (64, 64, 64), 64, 8, tt, [("one launch per mode", "", unf), ("fused", "", None)], a.rounds)
            ab_real("3d %s f32 128x128x32 K=64" % tt, (128, 128, 32), 64, 4, tt, [("one launch per mode", "", unf), ("fused", "", None)], a.rounds)
        return
    if a.which == "prof2":
        # one execute round of the shipped 128 x 128 fp32 tile kernels (c2c staged, r2c / c2r fused), for ncu
        ab("2d c2c f32 128x128 K=8192", (128, 128), 8192, 4, [("default", "", None)], 1)
        ab_real("2d r2c f32 128x128 K=8192", (128, 128), 8192, 4, "r2c", [("fused", "", None)], 1)
        ab_real("2d c2r f32 128x128 K=8192", (128, 128), 8192, 4, "c2r", [("fused", "", None)], 1)
        return
    if a.which == "prof":
        # three executes of each variant, for ncu
        x2 = {"BBFFT_CUDA_F32X2": "1"}
        ab("2d c2c f32 128x128 K=8192", (128, 128), 8192, 4, [("plain", "", None), ("x2", "", x2), ("SG=-1,BK=1", "SG=-1,BK=1", None)], 1)
        return
    if a.which == "x2":
        x2 = {"BBFFT_CUDA_F32X2": "1"}
        ab("2d c2c f32 128x128 K=8192", (128, 128), 8192, 4, [("plain", "", None), ("x2", "", x2), ("x2 TH=512", "TH=512", x2)], a.rounds)
        ab("2d c2c f32 64x64 K=32768", (64, 64), 32768, 4, [("plain", "", None), ("x2", "", x2)], a.rounds)
        for n in (343, 490, 225, 441, 512, 405, 256, 64, 7):
            ab("1d c2c f32 M=16 N=%d" % n, (n,), (1 << 30) // (16 * n * 8), 4, [("plain", "", None), ("x2", "", x2)], a.rounds, M=16)
        return
    sg = lambda rows: {"BBFFT_CUDA_TILE_STAGE": str(rows)}
    ab("2d c2c f32 128x128 K=8192", (128, 128), 8192, 4, [
        ("plain", "SG=0", None), ("default", "", None), ("PS=1", "PS=1,SG=0", None), ("SG=-1", "SG=-1", None), ("SG=64", "SG=64", None),
        ("SG=32", "SG=32", None), ("SG=-1,PADK=64", "SG=-1,PADK=64", None), ("SG=-1,TH=512", "SG=-1,TH=512", None),
        ("SG=-1,TH=512,RA=16x8,RB=16x8", "SG=-1,TH=512,RA=16x8,RB=16x8", None),
        ("TH=512", "TH=512", None), ("SG=-1,BK=1", "SG=-1,BK=1", None), ("SG=64,BK=1", "SG=64,BK=1", None),
        ("SG=-1,BK=1,TH=512", "SG=-1,BK=1,TH=512", None), ("SG=-1,BK=1,PADK=64", "SG=-1,BK=1,PADK=64", None),
    ], a.rounds)
    ab("2d c2c f64 64x64 K=16384", (64, 64), 16384, 8, [
        ("plain", "", None), ("SG=-1", "SG=-1", None), ("SG=-1,MB=1", "SG=-1,MB=1", None), ("SG=16", "SG=16", None),
        ("SG=32", "SG=32", None), ("SG=-1,BK=1", "SG=-1,BK=1", None), ("SG=-1,MB=1,BK=1", "SG=-1,MB=1,BK=1", None),
    ], a.rounds)
    ab("2d c2c f32 64x64 K=32768", (64, 64), 32768, 4, [
        ("plain", "", None), ("SG=-1", "SG=-1", None), ("SG=-1,MB=2", "SG=-1,MB=2", None), ("SG=-1,MB=3", "SG=-1,MB=3", None),
        ("SG=-1,MB=3,BK=1", "SG=-1,MB=3,BK=1", None),
    ], a.rounds)
    ab("2d c2c f64 128x64 K=8192", (128, 64), 8192, 8, [("plain", "SG=0", None), ("default", "", None), ("SG=-1,BK=0", "SG=-1,BK=0", None)], a.rounds)
    ab("2d c2c f32 256x64 K=8192", (256, 64), 8192, 4, [("plain", "SG=0", None), ("default", "", None), ("SG=-1,BK=0", "SG=-1,BK=0", None)], a.rounds)
    ab("3d c2c f64 64^3 K=64", (64, 64, 64), 64, 8, [
        ("plain", "", None), ("STAGE=-1", "", sg(-1)), ("STAGE=32", "", sg(32)), ("STAGE=16", "", sg(16)),
        ("STAGE=-1,BULK", "", dict(sg(-1), BBFFT_CUDA_TILE_BULK="1")),
    ], a.rounds)
    ab("3d c2c f64 64^3 K=256", (64, 64, 64), 256, 8, [("plain", "", None), ("STAGE=-1", "", sg(-1))], a.rounds)
    ab("2d c2c f32 128x128 K=64 (L2 resident)", (128, 128), 64, 4, [("plain", "", None), ("SG=-1", "SG=-1", None)], a.rounds)


if __name__ == "__main__":
    main()
