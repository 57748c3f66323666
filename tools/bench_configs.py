"""Bandwidth / GFLOP/s of the BASELINE.json configurations other than the headline sweep, with
cuFFT (through torch.fft) timed beside each on the same tensors, and a numerics check against
torch.fft in float64.

  C1  1d c2c fp32 N=64 M=1 K=16384 out-of-place
  C3  1d r2c / c2r fp32 N=256 M=1 K=2^20, in-place and out-of-place
  C4  3d c2c fp64 64x64x64 K=64 ; 2d c2c fp32 128x128 K=64
  C5  double-batched 1d c2c with load/store callbacks (identity callbacks: overhead), tft shapes
  R   r2c/c2r sweep (M=16, seven-smooth even+odd N), --real-sweep

Usage: python tools/bench_configs.py [--which c1,c3,c4,c5] [--real-sweep] [--out file.csv]
Algorithmic bytes follow benchmark/adapter.hpp:21-25,51-52.
"""
import argparse
import importlib
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("double-batched-fft-library_b200")
aot = importlib.import_module("double-batched-fft-library_b200.aot")

L2_FLUSH = None


def flush_l2():
    global L2_FLUSH
    if L2_FLUSH is None:
        L2_FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    L2_FLUSH.zero_()


def time_fn(fn, reps=10, inner=1, flush=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            flush_l2()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3 / inner)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def rel_l2(a, b):
    a = a.to(torch.complex128) if a.is_complex() else a.to(torch.float64)
    b = b.to(torch.complex128) if b.is_complex() else b.to(torch.float64)
    return float((a - b).norm() / b.norm())


def rdt(fp):
    return torch.float32 if fp == 4 else torch.float64


def cdt(fp):
    return torch.complex64 if fp == 4 else torch.complex128


def report(rows, name, fp, shape, nbytes, flops, t_best, t_med, t_cufft, err, kernels, note=""):
    row = dict(config=name, fp=fp, shape="x".join(map(str, shape)), time_us=t_best * 1e6, time_med_us=t_med * 1e6,
               GBs=nbytes / t_best * 1e-9, GFLOPs=flops / t_best * 1e-9,
               cufft_us=(t_cufft * 1e6 if t_cufft else None),
               cufft_GBs=(nbytes / t_cufft * 1e-9 if t_cufft else None),
               speedup_vs_cufft=(t_cufft / t_best if t_cufft else None), err=err, launches=len(kernels),
               kernel=kernels[0] if kernels else "", note=note)
    rows.append(row)
    print(json.dumps(row), flush=True)


def bench_c2c_1d(rows, name, fp, M, N, K, stream, callbacks=None, inplace=False, note=""):
    x = torch.view_as_complex(torch.rand(K, N, M, 2, dtype=rdt(fp), device="cuda"))
    y = x.clone() if inplace else torch.empty_like(x)
    cfg = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=inplace, callbacks=callbacks)
    plan = pkg.Plan(cfg, stream=stream)
    if inplace:
        plan.execute(y)
    else:
        plan.execute(x, y)
    torch.cuda.synchronize()
    kc = min(K, 32)
    err = rel_l2(y[:kc], torch.fft.fft(x[:kc].to(torch.complex128), dim=1))
    nbytes = 2.0 * M * N * K * 2 * fp
    small = nbytes < (200 << 20)
    inner = 20 if small else 2
    fn = (lambda: plan.execute(y)) if inplace else (lambda: plan.execute(x, y))
    tb, tm = time_fn(fn, inner=inner)
    tc, _ = time_fn(lambda: torch.fft.fft(x, dim=1, out=None), inner=inner)
    flops = 5.0 * N * math.log2(N) * M * K
    report(rows, name, fp, (M, N, K), nbytes, flops, tb, tm, tc, err, plan.kernel_names,
           note + (" L2-resident, back-to-back launches" if small else ""))
    plan.close()


def bench_c2c_1d_graph(rows, name, fp, M, N, K, stream, launches=20):
    """The same plan with `launches` executes captured in one CUDA graph: what a launch-bound caller
    (small, L2-resident batches) gets when it replays a graph instead of launching kernel by kernel."""
    x = torch.view_as_complex(torch.rand(K, N, M, 2, dtype=rdt(fp), device="cuda"))
    y = torch.empty_like(x)
    cfg = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
    plan = pkg.Plan(cfg, stream=stream)
    plan.execute(x, y)
    torch.cuda.synchronize()
    err = rel_l2(y[: min(K, 32)], torch.fft.fft(x[: min(K, 32)].to(torch.complex128), dim=1))
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(launches):
            plan.execute(x, y, stream=torch.cuda.current_stream().cuda_stream)
    tb, tm = time_fn(graph.replay, inner=1)
    nbytes = 2.0 * M * N * K * 2 * fp
    flops = 5.0 * N * math.log2(N) * M * K
    report(rows, name, fp, (M, N, K), nbytes, flops, tb / launches, tm / launches, None, err, plan.kernel_names,
           "per launch, %d launches replayed as one CUDA graph" % launches)
    plan.close()


def bench_real_1d(rows, name, fp, M, N, K, stream, ttype, inplace, cufft=True):
    nh = N // 2 + 1
    d = pkg.FORWARD if ttype == pkg.R2C else pkg.BACKWARD
    cfg = pkg.make_config(1, [M, N, K], fp, d, ttype, inplace=inplace)
    if inplace and pkg.describe(cfg)["inplace_unsupported"]:
        return
    plan = pkg.Plan(cfg, stream=stream)
    nbytes = float(N * fp + nh * 2 * fp) * M * K
    flops = 2.5 * N * math.log2(N) * M * K
    if ttype == pkg.R2C:
        nrow = 2 * nh if inplace else N
        buf = torch.rand(K, nrow, M, dtype=rdt(fp), device="cuda")
        xs = buf[:, :N, :]
        want = torch.fft.rfft(xs[: min(K, 16)].to(torch.float64), dim=1)
        if inplace:
            work = buf.clone()
            plan.execute(work)
            torch.cuda.synchronize()
            # the spectrum overlays the padded real rows: 2*nh*M reals per k = nh*M complex
            got = torch.view_as_complex(work.view(K, nh, M, 2))
            fn = lambda: plan.execute(work)
        else:
            out = torch.empty(K, nh, M, dtype=cdt(fp), device="cuda")
            plan.execute(buf, out)
            torch.cuda.synchronize()
            got = out
            fn = lambda: plan.execute(buf, out)
        err = rel_l2(got[: min(K, 16)], want)
        tc = time_fn(lambda: torch.fft.rfft(xs, dim=1), inner=2)[0] if cufft else None
    else:
        spec = torch.fft.rfft(torch.rand(K, N, M, dtype=rdt(fp), device="cuda"), dim=1).contiguous()
        want = torch.fft.irfft(spec[: min(K, 16)].to(torch.complex128), n=N, dim=1) * N
        if inplace:
            work = spec.clone()
            plan.execute(work)
            torch.cuda.synchronize()
            got = torch.view_as_real(work).view(K, -1)[:, : 2 * nh * M].view(K, 2 * nh, M)[:, :N, :]
            err = rel_l2(got[: min(K, 16)], want)
            fn = lambda: plan.execute(work)
        else:
            out = torch.empty(K, N, M, dtype=rdt(fp), device="cuda")
            plan.execute(spec, out)
            torch.cuda.synchronize()
            err = rel_l2(out[: min(K, 16)], want)
            fn = lambda: plan.execute(spec, out)
        tc = time_fn(lambda: torch.fft.irfft(spec, n=N, dim=1), inner=2)[0] if cufft else None
    tb, tm = time_fn(fn, inner=2)
    report(rows, name, fp, (M, N, K), nbytes, flops, tb, tm, tc, err, plan.kernel_names,
           ("in-place" if inplace else "out-of-place"))
    plan.close()


def bench_nd(rows, name, fp, dims, K, stream, tune="", env=None):
    """`env`: BBFFT_CUDA_ND_* switches in force while the plan is created (chain / fuse / blocking)."""
    shape = [1] + list(dims) + [K]
    cfg = pkg.make_config(len(dims), shape, fp, pkg.FORWARD, pkg.C2C, inplace=False)
    saved = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        plan = pkg.Plan(cfg, stream=stream, tune=tune)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    tdims = [K] + list(reversed(dims))
    x = torch.view_as_complex(torch.rand(*tdims, 2, dtype=rdt(fp), device="cuda"))
    y = torch.empty_like(x)
    plan.execute(x, y)
    torch.cuda.synchronize()
    axes = tuple(range(1, len(dims) + 1))
    kc = min(K, 4)
    err = rel_l2(y[:kc], torch.fft.fftn(x[:kc].to(torch.complex128), dim=axes))
    n = 1
    for d in dims:
        n *= d
    nbytes = 2.0 * 2 * fp * n * K
    flops = 5.0 * n * math.log2(n) * K
    small = nbytes < (200 << 20)
    inner = 20 if small else 2
    tb, tm = time_fn(lambda: plan.execute(x, y), inner=inner)
    tc, _ = time_fn(lambda: torch.fft.fftn(x, dim=axes), inner=inner)
    report(rows, name, fp, shape, nbytes, flops, tb, tm, tc, err, plan.kernel_names,
           "L2-resident, back-to-back launches" if small else "")
    plan.close()


def bench_nd_real(rows, name, fp, dims, K, stream, ttype, env=None):
    """r2c / c2r 2d / 3d, M = 1, out of place; algorithmic bytes = real tensor + spectrum tensor
    (SURVEY 8d generalised: N reals in, (N1/2+1) N2 .. complex out); cuFFT = torch.fft.rfftn / irfftn."""
    fwd = ttype == pkg.R2C
    cfg = pkg.make_config(len(dims), [1] + list(dims) + [K], fp, pkg.FORWARD if fwd else pkg.BACKWARD, ttype, inplace=False)
    saved = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        plan = pkg.Plan(cfg, stream=stream)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    tdims = [K] + list(reversed(dims))
    axes = tuple(range(1, len(dims) + 1))
    x = torch.rand(*tdims, dtype=rdt(fp), device="cuda")
    kc = min(K, 4)
    n = 1
    for d in dims:
        n *= d
    nspec = (dims[0] // 2 + 1) * (n // dims[0])
    if fwd:
        y = torch.empty(tdims[:-1] + [dims[0] // 2 + 1], dtype=cdt(fp), device="cuda")
        plan.execute(x, y)
        torch.cuda.synchronize()
        err = rel_l2(y[:kc], torch.fft.rfftn(x[:kc].to(torch.float64), dim=axes))
        fn = lambda: plan.execute(x, y)
        cf = lambda: torch.fft.rfftn(x, dim=axes)
    else:
        spec = torch.fft.rfftn(x, dim=axes).contiguous()
        y = torch.empty_like(x)
        plan.execute(spec, y)
        torch.cuda.synchronize()
        err = rel_l2(y[:kc], torch.fft.irfftn(spec[:kc].to(torch.complex128), s=tdims[1:], dim=axes) * n)
        fn = lambda: plan.execute(spec, y)
        cf = lambda: torch.fft.irfftn(spec, s=tdims[1:], dim=axes)
    nbytes = float(n * fp + nspec * 2 * fp) * K
    flops = 2.5 * n * math.log2(n) * K
    tb, tm = time_fn(fn, inner=2)
    tc, _ = time_fn(cf, inner=2)
    report(rows, name, fp, [1] + list(dims) + [K], nbytes, flops, tb, tm, tc, err, plan.kernel_names)
    plan.close()


IDENTITY_CB = """
__device__ %(v)s load(%(v)s const* in, size_t offset) { return in[offset]; }
__device__ void store(%(v)s* out, size_t offset, %(v)s value) { out[offset] = value; }
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c1,c3,c4,c5")
    ap.add_argument("--real-sweep", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    which = set(args.which.split(","))
    stream = torch.cuda.current_stream().cuda_stream
    rows = []
    if "c1" in which:
        bench_c2c_1d(rows, "C1", 4, 1, 64, 16384, stream)
        bench_c2c_1d(rows, "C1-big", 4, 1, 64, (1 << 30) // (64 * 8), stream, note="same shape, K sized to 1 GiB")
        bench_c2c_1d(rows, "C1-big-f64", 8, 1, 64, (1 << 30) // (64 * 16), stream, note="same shape, K sized to 1 GiB")
    if "c3" in which:
        for ttype, nm in ((pkg.R2C, "C3-r2c"), (pkg.C2R, "C3-c2r")):
            for inplace in (False, True):
                bench_real_1d(rows, nm, 4, 1, 256, 1 << 20, stream, ttype, inplace)
        for ttype, nm in ((pkg.R2C, "C3-r2c-M16"), (pkg.C2R, "C3-c2r-M16")):
            for inplace in (False, True):
                bench_real_1d(rows, nm, 4, 16, 256, 1 << 16, stream, ttype, inplace)
    if "c4" in which:
        bench_nd(rows, "C4-3d", 8, (64, 64, 64), 64, stream)
        bench_nd(rows, "C4-3d-chained", 8, (64, 64, 64), 64, stream, env={"BBFFT_CUDA_ND_CHAIN": "1"})
        bench_nd(rows, "C4-3d-multipass", 8, (64, 64, 64), 64, stream, env={"BBFFT_CUDA_ND_FUSE": "0"})
        if "chain" in which:
            for kb in (1, 2, 8):
                bench_nd(rows, "C4-3d-chain-kblock%d" % kb, 8, (64, 64, 64), 64, stream,
                         env={"BBFFT_CUDA_ND_CHAIN": "1", "BBFFT_CUDA_ND_CHAIN_KBLOCK": str(kb)})
            bench_nd(rows, "C4-3d-chain-1cta", 8, (64, 64, 64), 64, stream, env={"BBFFT_CUDA_ND_CHAIN": "1", "BBFFT_CUDA_ND_CHAIN_CTAS": "1"})
            bench_nd(rows, "C4-3d-f32-256^3-chained", 4, (256, 256, 256), 8, stream, env={"BBFFT_CUDA_ND_CHAIN": "1"})
            bench_nd(rows, "C4-3d-f32-256^3", 4, (256, 256, 256), 8, stream)
        bench_nd(rows, "C4-2d", 4, (128, 128), 64, stream)
        bench_nd(rows, "C4-2d-big", 4, (128, 128), 8192, stream)
        bench_nd(rows, "C4-2d-big-multipass", 4, (128, 128), 8192, stream, env={"BBFFT_CUDA_ND_FUSE": "0"})
        if "chain" in which:
            bench_nd(rows, "C4-2d-big-chained-passes", 4, (128, 128), 8192, stream, env={"BBFFT_CUDA_ND_FUSE": "0", "BBFFT_CUDA_ND_CHAIN": "1"})
            bench_nd(rows, "C4-2d-chained-passes", 4, (128, 128), 64, stream, env={"BBFFT_CUDA_ND_FUSE": "0", "BBFFT_CUDA_ND_CHAIN": "1"})
    if "c5" in which:
        for n in (64, 256):
            k = (1 << 30) // (16 * n * 8)
            bench_c2c_1d(rows, "C5-plain", 4, 16, n, k, stream)
            bench_c2c_1d(rows, "C5-identity-cb", 4, 16, n, k, stream,
                         callbacks=(IDENTITY_CB % dict(v="float2"), "load", "store", "cuda"))
        for (m, n) in ((16, 16), (1120, 32), (128, 512), (70, 16)):
            k = max(1, int(512e6) // (16 * m * n))
            bench_c2c_1d(rows, "C5-tft-shape", 8, m, n, k, stream, note="tft.cpp shape")
    if "c1" in which:
        try:  # last: a failed stream capture must not cost the rows above
            bench_c2c_1d_graph(rows, "C1-graph", 4, 1, 64, 16384, stream)
        except Exception as ex:
            print("C1-graph failed:", str(ex)[:200], file=sys.stderr)
    if args.real_sweep:
        sizes = [n for n in aot.smooth_sizes() if n in (2, 3, 4, 7, 8, 15, 16, 27, 32, 49, 64, 100, 105, 128, 135,
                                                         200, 243, 256, 315, 343, 384, 400, 441, 480, 500, 512)]
        for fp in (4, 8):
            for n in sizes:
                k = max(2, (1 << 30) // (16 * n * fp)) // 2 * 2
                for ttype, nm in ((pkg.R2C, "R-r2c"), (pkg.C2R, "R-c2r")):
                    bench_real_1d(rows, nm, fp, 16, n, k, stream, ttype, False, cufft=False)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            keys = list(rows[0].keys())
            f.write(",".join(keys) + "\n")
            for r in rows:
                f.write(",".join(str(r[k]) for k in keys) + "\n")


if __name__ == "__main__":
    main()
