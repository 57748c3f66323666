#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 batched small-FFT path.

Workload (BASELINE.json configs[1]): 1d c2c, fp32 + fp64, double-batched M=16, N = the 105
seven-smooth sizes in [2,512] (the reference's own sweep, tools/powers.py -p 4 2 512), K sized so
that every input tensor is ~1 GiB (benchmark/test.hpp:18-30), out-of-place, forward.  One "step"
is one execute of every (precision, N) plan of the sweep: 210 kernel launches, ~210 GiB of
input streamed.  Metric: aggregate GFLOP/s under the reference's 5*N*log2(N) convention
(benchmark/adapter.hpp:51) plus the achieved HBM GB/s (2*N*sizeof(complex) bytes per transform,
benchmark/adapter.hpp:52).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1: launched by torchrun, one rank per GPU; every rank owns its own contiguous K slab of the
same size (weak scaling), no collective on the data path; the timed
region is bracketed by barriers and the max over ranks is reported.
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "double-batched-fft-library_b200"

M_BATCH = 16
TENSOR_BYTES = 1 << 30
# dram__bytes_read.sum + dram__bytes_write.sum of one launch (2 GiB = 2147483648 algorithmic bytes)
# from the ncu --set full capture of the N=64 fp32 kernel of this sweep
NCU_TRAFFIC_PER_LAUNCH = 1073752000 + 1022798000
NCU_TRAFFIC_SOURCE = ("profiles/r02j_ncu_full_summary.txt (c2c f32 M=16 N=64: read 1.073752 GB + write 1.022798 GB; "
                      "f64 N=490: 1.073669 + 1.026042 GB; f64 N=486: 1.073745 + 1.022654 GB; f32 N=225: 1.073906 + 1.022284 GB; "
                      "the tail of the output is still dirty in L2 when the kernel ends)")


def sweep_sizes():
    aot = importlib.import_module(PKG + ".aot")
    return aot.smooth_sizes()


def flops_c2c(n, transforms):
    return 5.0 * n * math.log2(n) * transforms


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's generated kernels under the host emulator (oracle/_ref), all cores
# ------------------------------------------------------------------------------------------------
REF_SAMPLE = [(4, 2, 4096), (4, 8, 2048), (4, 27, 512), (4, 64, 384), (4, 105, 192), (4, 256, 96), (4, 512, 48),
              (8, 8, 2048), (8, 64, 384), (8, 105, 192), (8, 512, 48)]


def run_reference_sample(steps, warmup, threads):
    """Returns (gflops, gbs, seconds_per_step, kind, sample description)."""
    import numpy as np
    from oracle import oracle, refemu
    kind = "reference" if refemu.available() else "port"
    rng = np.random.default_rng(0)
    work = []
    for fp, n, k in REF_SAMPLE:
        dt = np.complex64 if fp == 4 else np.complex128
        x = (rng.standard_normal((k, n, M_BATCH)) + 1j * rng.standard_normal((k, n, M_BATCH))).astype(dt)
        y = np.empty_like(x)
        if kind == "reference":
            refemu.set_threads(threads)
            plan = refemu.Plan(refemu.make_config(1, [M_BATCH, n, k], fp, -1, 0, inplace=False))
            fn = (lambda p=plan, a=x, b=y: p.execute(a, b))
        else:
            cfg = oracle.make_config(1, [M_BATCH, n, k], fp, -1, 0, inplace=False)
            fn = (lambda c=cfg, a=x, b=y: oracle.bbfft(c, a, b))
        work.append((fn, flops_c2c(n, M_BATCH * k), 2.0 * M_BATCH * n * k * 2 * fp))
    for _ in range(warmup):
        for fn, _, _ in work:
            fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        for fn, _, _ in work:
            fn()
    dt = (time.perf_counter() - t0) / steps
    fl = sum(w[1] for w in work)
    by = sum(w[2] for w in work)
    sample = "c2c M=16 (fp,N,K) in %s: the sweep's shapes with K cut to a CPU-sized batch" % (REF_SAMPLE,)
    return fl / dt * 1e-9, by / dt * 1e-9, dt, kind, sample


POCKETFFT_TENSOR_BYTES = 128 << 20  # well beyond the host's last-level cache


def run_pocketfft_sample(steps, warmup, threads):
    """An honest CPU line: scipy.fft (pocketfft, C++) over ALL 210 (precision, N) shapes of the sweep,
    same M=16 double-batched layout (transform along axis 1 of [K][N][M]), K cut so that every tensor is
    128 MiB (out of cache), `threads` workers.  Not the reference's code -- the reference has no CPU implementation of
    this path that runs here -- but the best CPU FFT in this image."""
    import numpy as np
    import scipy.fft
    rng = np.random.default_rng(0)
    work = []
    for fp in (4, 8):
        dt = np.complex64 if fp == 4 else np.complex128
        base = (rng.standard_normal(POCKETFFT_TENSOR_BYTES // (2 * fp)) +
                1j * rng.standard_normal(POCKETFFT_TENSOR_BYTES // (2 * fp))).astype(dt)
        for n in sweep_sizes():
            k = max(1, POCKETFFT_TENSOR_BYTES // (M_BATCH * n * 2 * fp))
            x = base[: k * n * M_BATCH].reshape(k, n, M_BATCH)
            work.append((x, flops_c2c(n, M_BATCH * k), 2.0 * M_BATCH * n * k * 2 * fp))
    def step():
        for x, _, _ in work:
            scipy.fft.fft(x, axis=1, workers=threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    fl = sum(w[1] for w in work)
    by = sum(w[2] for w in work)
    return {"value": fl / dt * 1e-9, "unit": "GFLOP/s", "cores": threads, "kind": "pocketfft", "gbs": by / dt * 1e-9,
            "seconds_per_step": dt,
            "sample": "scipy.fft.fft(x, axis=1, workers=%d) on all 210 (fp, N) shapes of the sweep, M=16, K cut to "
                      "128 MiB per tensor (out-of-place, complex input resident in host memory)" % threads}


def reference_arm(args, emit=print):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, args.steps)
    warmup = max(3, args.warmup)  # same warm-up rule as the repo arm
    gf, gbs, dt, kind, sample = run_reference_sample(steps, warmup, threads)
    try:
        pocket = run_pocketfft_sample(2, 1, threads)
    except Exception as ex:
        pocket = {"value": None, "kind": "pocketfft", "error": str(ex)[:200]}
    line = {
        "impl": "reference",
        "metric": "c2c GFLOP/s (5N*log2N), 1d double-batched sweep N=2..512",
        "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "gbs": gbs,
        "what_this_is": ("the reference's generated OpenCL-C kernels run work-item by work-item under a host emulator "
                         "(oracle/_ref; the reference's real backends need an OpenCL/SYCL/Level-Zero device, SURVEY.md 8c) "
                         "on an 11-shape sample of the sweep with K cut to CPU size: a parity-grade stand-in, NOT a tuned "
                         "CPU FFT.  `pocketfft` beside it is the honest CPU number for the same layout."),
        "same_config": False,
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": threads, "kind": kind, "sample": sample,
                         "pocketfft": pocket},
        "pocketfft": pocket,
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


def workload_config(gpus):
    return {"workload": "1d c2c fp32+fp64 sweep, 105 seven-smooth N in [2,512], M=16, K=1GiB/(16*N*sizeof(complex)) per GPU, out-of-place, forward",
            "M": M_BATCH, "tensor_bytes_per_gpu": TENSOR_BYTES, "n_sizes": len(sweep_sizes()),
            "l2": "inputs_larger_than_l2 (2 GiB streamed per launch)", "parallelism": "k-sharded x%d, no collective" % gpus}


def measure_copy_ceiling(torch, dist, world, dev, barrier, nbytes, reps=3):
    """What the host link gives this job with no FFT at all: H2D of `nbytes` on one stream while D2H of
    `nbytes` runs on another (pinned buffers, all ranks at once), best of `reps`.  GB/s counts both
    directions, whole job."""
    hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dout = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = None
    for _ in range(reps + 1):
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        best = dt if best is None else min(best, dt)
    return 2.0 * nbytes * world / best * 1e-9


def measure_e2e(args, torch, dist, plans, world, dev, barrier, total_flops):
    """The same sweep through bbfft_cuda_plan_execute_host with pinned HOST buffers: every plan's
    input is copied host->device, transformed and copied back inside the timed region."""
    if args.e2e_steps <= 0:
        return None
    hin = {4: torch.empty(TENSOR_BYTES // 4, dtype=torch.float32).pin_memory(),
           8: torch.empty(TENSOR_BYTES // 8, dtype=torch.float64).pin_memory()}
    hout = torch.empty(TENSOR_BYTES, dtype=torch.uint8).pin_memory()
    for t in hin.values():
        t.uniform_(0.0, 1.0)
    h2d = d2h = 0
    for fp, n, k, plan in plans:
        nb = M_BATCH * n * k * 2 * fp
        h2d += nb
        d2h += nb

    def e2e_step():
        for fp, n, k, plan in plans:
            nb = M_BATCH * n * k * 2 * fp
            plan.execute_host(hin[fp][: nb // fp], hout[:nb])

    # untimed: the first host call creates the device ring and its streams
    fp0, n0, k0, plan0 = plans[0]
    plan0.execute_host(hin[fp0][: M_BATCH * n0 * k0 * 2], hout[: M_BATCH * n0 * k0 * 2 * fp0])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    t_e2e = (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
        t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    gbs = (h2d + d2h) * world / t_e2e * 1e-9
    try:
        ceiling = measure_copy_ceiling(torch, dist, world, dev, barrier, TENSOR_BYTES)
    except Exception:
        ceiling = None
    # whole-job figures, like `value`: every rank copies its own slab in and out
    return {"value": total_flops / t_e2e * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d * world,
            "d2h_bytes_per_step": d2h * world, "ms_per_step": t_e2e * 1e3, "steps": args.e2e_steps,
            "gbs": gbs, "copy_ceiling_gbs": ceiling,
            "frac_of_copy_ceiling": (gbs / ceiling if ceiling else None),
            "copy_ceiling": "pinned 1 GiB H2D || 1 GiB D2H on two streams, every rank at once, no FFT (same units as gbs)",
            "api": "bbfft_cuda_plan_execute_host: pinned host buffers; slabs of 16 MiB go H2D -> kernel -> D2H over a "
                   "4-slot device ring on three streams"}


def other_configs(peak):
    """BASELINE.json configs 1, 3, 4, 5 (single launches, outside the timed region of the headline
    sweep): algorithmic GB/s, fraction of the measured HBM peak, cuFFT (torch.fft) on the same
    tensors.  Measured by tools/bench_configs.py."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import contextlib
    import io
    import torch
    import bench_configs as bc
    stream = torch.cuda.current_stream().cuda_stream
    rows = []
    with contextlib.redirect_stdout(io.StringIO()):
        bc.bench_c2c_1d(rows, "C1 1d c2c f32 N=64 M=1 K=16384 (L2 resident)", 4, 1, 64, 16384, stream)
        bc.bench_c2c_1d(rows, "C1 shape at 1 GiB", 4, 1, 64, (1 << 30) // (64 * 8), stream)
        for ttype, nm in ((bc.pkg.R2C, "C3 r2c f32 N=256 K=2^20"), (bc.pkg.C2R, "C3 c2r f32 N=256 K=2^20")):
            for inplace in (False, True):
                bc.bench_real_1d(rows, nm, 4, 1, 256, 1 << 20, stream, ttype, inplace)
        bc.bench_nd(rows, "C4 3d c2c f64 64^3 K=64", 8, (64, 64, 64), 64, stream)
        bc.bench_nd(rows, "C4 3d c2c f64 64^3 K=64, chained (one persistent launch, L2-resident intermediate)", 8,
                    (64, 64, 64), 64, stream, env={"BBFFT_CUDA_ND_CHAIN": "1"})
        bc.bench_nd(rows, "C4 3d c2c f64 64^3 K=64, multi-pass (reference decomposition)", 8, (64, 64, 64), 64, stream,
                    env={"BBFFT_CUDA_ND_FUSE": "0"})
        bc.bench_nd(rows, "C4 2d c2c f32 128^2 K=64 (L2 resident)", 4, (128, 128), 64, stream)
        bc.bench_nd(rows, "C4 2d shape at 1 GiB", 4, (128, 128), 8192, stream)
        # real nd transforms: modes 1 and 2 fused in one tile kernel, against one launch per mode (the reference's
        # nd_fft decomposition)
        for ttype, nm in ((bc.pkg.R2C, "r2c"), (bc.pkg.C2R, "c2r")):
            bc.bench_nd_real(rows, "nd real: 2d %s f32 128^2 K=8192, fused tile" % nm, 4, (128, 128), 8192, stream, ttype)
            bc.bench_nd_real(rows, "nd real: 2d %s f32 128^2 K=8192, one launch per mode" % nm, 4, (128, 128), 8192, stream,
                             ttype, env={"BBFFT_CUDA_ND_FUSE_REAL": "0"})
            bc.bench_nd_real(rows, "nd real: 3d %s f64 64^3 K=64, fused tile + 1d pass" % nm, 8, (64, 64, 64), 64, stream, ttype)
            bc.bench_nd_real(rows, "nd real: 3d %s f64 64^3 K=64, one launch per mode" % nm, 8, (64, 64, 64), 64, stream, ttype,
                             env={"BBFFT_CUDA_ND_FUSE_REAL": "0"})
        bc.bench_c2c_1d(rows, "C5 c2c f32 M=16 N=256 identity load/store callbacks", 4, 16, 256, (1 << 30) // (16 * 256 * 8),
                        stream, callbacks=(bc.IDENTITY_CB % dict(v="float2"), "load", "store", "cuda"))
        try:  # last, and optional: a failed stream capture must not cost the rows above
            bc.bench_c2c_1d_graph(rows, "C1 1d c2c f32 N=64 M=1 K=16384 (L2 resident)", 4, 1, 64, 16384, stream)
        except Exception:
            pass
    out = []
    for r in rows:
        out.append({"config": r["config"] + (" " + r["note"] if r["note"] else ""), "GBs": round(r["GBs"], 1),
                    "frac_of_peak": round(r["GBs"] / peak, 4), "GFLOPs": round(r["GFLOPs"], 1),
                    "time_us": round(r["time_us"], 2), "cufft_GBs": round(r["cufft_GBs"], 1) if r["cufft_GBs"] else None,
                    "rel_l2_vs_fp64": r["err"], "launches": r["launches"]})
    return out


def sharded_configs(torch, dist, pkg, world, rank, dev, stream, peak):
    """BASELINE configs 3 and 5 and a strong-scaling point, timed on every rank of an N-GPU run (no
    collective on the data path; the per-row time is the max over ranks, CUDA events on the launching
    stream, median of 9 with the L2 flushed before every launch).

    weak rows    -- every rank runs the full per-GPU problem of the config (C3: r2c/c2r fp32 N=256 M=1
                    K=2^20 in/out of place; C5: c2c fp32 M=16 N=256 with identity load/store callbacks and
                    one transpose_fft_transpose shape), aggregate GB/s = world * bytes / t.
    strong rows  -- ONE ~1 GiB c2c tensor (the headline's per-GPU problem) split into `world` contiguous k
                    slabs (128 MiB per GPU at world = 8, the smallest slab SURVEY.md 8e allows); rank 0 also
                    times the whole tensor on its own GPU, so efficiency = t_whole / (world * t_slab)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as bc

    def timed(fn, reps=9):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            bc.flush_l2()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        ts.sort()
        return ts[len(ts) // 2]

    def reduce_max(t):
        if world == 1:
            return t
        v = torch.tensor([t], device=dev, dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    rows = []

    def weak_row(name, nbytes, flops, fn, launches=1):
        if world > 1:
            dist.barrier()
        t = reduce_max(timed(fn))
        rows.append({"config": name, "scaling": "weak", "time_us": round(t * 1e6, 2), "launches": launches,
                     "GBs": round(nbytes * world / t * 1e-9, 1), "GBs_per_gpu_frac_of_peak": round(nbytes / t * 1e-9 / peak, 4),
                     "GFLOPs": round(flops * world / t * 1e-9, 1)})

    # ---- C3: r2c / c2r fp32 N=256 M=1 K=2^20, in- and out-of-place
    N, K = 256, 1 << 20
    nh = N // 2 + 1
    nbytes = float(N * 4 + nh * 8) * K
    flops = 2.5 * N * math.log2(N) * K
    xr = torch.rand(K, N, dtype=torch.float32, device=dev)
    spec = torch.empty(K, nh, dtype=torch.complex64, device=dev)
    back = torch.empty(K, N, dtype=torch.float32, device=dev)
    pad = torch.zeros(K, 2 * nh, dtype=torch.float32, device=dev)
    for ttype, nm, src, dst in ((pkg.R2C, "r2c", xr, spec), (pkg.C2R, "c2r", spec, back)):
        d = pkg.FORWARD if ttype == pkg.R2C else pkg.BACKWARD
        plan = pkg.Plan(pkg.make_config(1, [1, N, K], 4, d, ttype, inplace=False), stream=stream, device=dev.index)
        weak_row("C3 %s f32 N=256 M=1 K=2^20 out-of-place" % nm, nbytes, flops, lambda: plan.execute(src, dst))
        plan.close()
        plan = pkg.Plan(pkg.make_config(1, [1, N, K], 4, d, ttype, inplace=True), stream=stream, device=dev.index)
        weak_row("C3 %s f32 N=256 M=1 K=2^20 in-place" % nm, nbytes, flops, lambda: plan.execute(pad))
        plan.close()
    del xr, spec, back, pad

    # ---- C5: callbacks (identity load/store: the cost of the hooks) and a transpose_fft_transpose shape
    n5 = 256
    k5 = TENSOR_BYTES // (M_BATCH * n5 * 8)
    x5 = torch.view_as_complex(torch.rand(k5, n5, M_BATCH, 2, dtype=torch.float32, device=dev))
    y5 = torch.empty_like(x5)
    cb = (bc.IDENTITY_CB % dict(v="float2"), "load", "store", "cuda")
    for label, callbacks in (("plain", None), ("identity load/store callbacks", cb)):
        plan = pkg.Plan(pkg.make_config(1, [M_BATCH, n5, k5], 4, pkg.FORWARD, pkg.C2C, inplace=False, callbacks=callbacks),
                        stream=stream, device=dev.index)
        weak_row("C5 c2c f32 M=16 N=256 K=%d %s" % (k5, label), 2.0 * x5.numel() * 8,
                 5.0 * n5 * math.log2(n5) * M_BATCH * k5, lambda: plan.execute(x5, y5))
        plan.close()
    del x5, y5
    mt, nt = 128, 512  # examples/transpose_fft_transpose/tft.cpp:20,100-103 (fp64)
    kt = max(1, int(512e6) // (16 * mt * nt))
    xt = torch.view_as_complex(torch.rand(kt, nt, mt, 2, dtype=torch.float64, device=dev))
    yt = torch.empty_like(xt)
    plan = pkg.Plan(pkg.make_config(1, [mt, nt, kt], 8, pkg.FORWARD, pkg.C2C, inplace=False), stream=stream, device=dev.index)
    weak_row("C5 tft shape c2c f64 M=128 N=512 K=%d (double-batched plan, no transposes)" % kt, 2.0 * xt.numel() * 16,
             5.0 * nt * math.log2(nt) * mt * kt, lambda: plan.execute(xt, yt))
    plan.close()
    del xt, yt

    # ---- strong scaling: one 1 GiB tensor split into `world` k slabs
    for fp, n in ((4, 64), (4, 256), (4, 512), (8, 64), (8, 490)):
        k_total = TENSOR_BYTES // (M_BATCH * n * 2 * fp)
        k_rank = k_total // world
        rdt = torch.float32 if fp == 4 else torch.float64
        x = torch.view_as_complex(torch.rand(k_total if rank == 0 else k_rank, n, M_BATCH, 2, dtype=rdt, device=dev))
        y = torch.empty_like(x)
        slab = pkg.Plan(pkg.make_config(1, [M_BATCH, n, k_rank], fp, pkg.FORWARD, pkg.C2C, inplace=False), stream=stream,
                        device=dev.index)
        if world > 1:
            dist.barrier()
        t_slab = reduce_max(timed(lambda: slab.execute(x, y)))
        slab.close()
        t_whole = None
        if world > 1 and rank == 0:
            whole = pkg.Plan(pkg.make_config(1, [M_BATCH, n, k_total], fp, pkg.FORWARD, pkg.C2C, inplace=False), stream=stream,
                             device=dev.index)
            t_whole = timed(lambda: whole.execute(x, y))
            whole.close()
        if world > 1:
            dist.barrier()
        nb = 2.0 * M_BATCH * n * k_rank * world * 2 * fp
        row = {"config": "strong: c2c %s M=16 N=%d, K=%d total = %d per GPU (%.0f MiB in per GPU)" %
                         ("f32" if fp == 4 else "f64", n, k_rank * world, k_rank, nb / world / 2 / (1 << 20)),
               "scaling": "strong", "time_us": round(t_slab * 1e6, 2), "launches": 1,
               "GBs": round(nb / t_slab * 1e-9, 1), "GBs_per_gpu_frac_of_peak": round(nb / world / t_slab * 1e-9 / peak, 4)}
        if t_whole:
            row["one_gpu_time_us"] = round(t_whole * 1e6, 2)
            row["strong_efficiency"] = round(t_whole / (world * t_slab), 4)
            # what is not bandwidth: the fixed cost of a launch (measured ~5 us back to back) in a slab that
            # takes t_slab in total
            row["launch_overhead_share"] = round(5e-6 / t_slab, 4)
        rows.append(row)
        del x, y
    return rows


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--per-size", default="", help="write the per-size table to this CSV")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (C1, C3, C4, C5)")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL prints its version
    # with NCCL_DEBUG >= VERSION) are sent to stderr for the whole run
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(line):
        real_stdout.write(line + "\n")
        real_stdout.flush()

    if args.impl == "reference":
        reference_arm(args, emit)
        return

    import torch
    import torch.distributed as dist

    pkg = importlib.import_module(PKG)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)

    stream = torch.cuda.current_stream().cuda_stream
    sizes = sweep_sizes()
    # resident synthetic inputs: one 1 GiB input + 1 GiB output buffer per precision
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    bufs = {}
    for fp, rdt in ((4, torch.float32), (8, torch.float64)):
        n_real = TENSOR_BYTES // fp
        x = torch.empty(n_real, dtype=rdt, device=dev)
        x.uniform_(0.0, 1.0, generator=gen)
        y = torch.empty(n_real, dtype=rdt, device=dev)
        bufs[fp] = (x, y)
    plans = []
    cache = pkg.Cache()
    for fp in (4, 8):
        for n in sizes:
            k = max(1, TENSOR_BYTES // (M_BATCH * n * 2 * fp))
            cfg = pkg.make_config(1, [M_BATCH, n, k], fp, pkg.FORWARD, pkg.C2C, inplace=False)
            plan = pkg.Plan(cfg, stream=stream, device=local_rank, cache=cache)
            plans.append((fp, n, k, plan))
    launches_per_step = sum(p.launches_per_execute for _, _, _, p in plans)

    def run_step(events=None):
        for i, (fp, n, k, plan) in enumerate(plans):
            x, y = bufs[fp]
            if events is not None:
                events[i][0].record()
            plan.execute(x, y)
            if events is not None:
                events[i][1].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        run_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    per = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in plans]
           for _ in range(steps)]
    barrier()
    ev0.record()
    for s in range(steps):
        run_step(per[s])
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed = ev0.elapsed_time(ev1) * 1e-3
    if world > 1:
        t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())

    total_flops = sum(flops_c2c(n, M_BATCH * k) for _, n, k, _ in plans) * world
    total_bytes = sum(2.0 * M_BATCH * n * k * 2 * fp for fp, n, k, _ in plans) * world
    per_step = elapsed / steps
    value = total_flops / per_step * 1e-9
    gbs = total_bytes / per_step * 1e-9

    # per-launch kernel durations (CUDA events on the launching stream), this rank
    rows = []
    kernel_time = 0.0
    for i, (fp, n, k, plan) in enumerate(plans):
        ts = sorted(per[s][i][0].elapsed_time(per[s][i][1]) * 1e-3 for s in range(steps))
        avg = sum(ts) / len(ts)
        med = ts[len(ts) // 2]
        kernel_time += avg  # roofline.achieved: algorithmic bytes / AVERAGE launch duration
        b = 2.0 * M_BATCH * n * k * 2 * fp
        # the per-size table reports the median: one hiccup in one of the timed steps (seen once: 2.6 ms on
        # a 0.33 ms launch, profiles/r01l_per_size.csv fp64 N=63) should not brand a size as slow
        rows.append((fp, n, k, med, b / med * 1e-9, flops_c2c(n, M_BATCH * k) / med * 1e-9, plan.kernel_names[0], avg))
    peak, peak_src = peak_hbm()
    fracs = sorted(r[4] / peak for r in rows)
    worst = min(rows, key=lambda r: r[4])
    bytes_rank = total_bytes / world
    roofline = {
        "bound": "hbm", "achieved": bytes_rank / kernel_time * 1e-9, "peak": peak, "unit": "GB/s",
        "frac": bytes_rank / kernel_time * 1e-9 / peak, "traffic": NCU_TRAFFIC_PER_LAUNCH, "peak_source": peak_src,
        "traffic_source": NCU_TRAFFIC_SOURCE,
        "kernel": "bbk::fft1d<C> (all 210 instantiations of the sweep, per-launch CUDA events)",
        "algorithmic_bytes_per_launch": "2*N*sizeof(complex)*M*K = 2 GiB",
        "per_size_frac": {"min": fracs[0], "median": fracs[len(fracs) // 2], "max": fracs[-1],
                          "n_below_0.8": sum(1 for f in fracs if f < 0.8), "n_below_0.85": sum(1 for f in fracs if f < 0.85),
                          "statistic": "median over the timed steps"},
        "worst": {"fp": worst[0], "N": worst[1], "GBs": worst[4], "kernel": worst[6]},
        "below_0.8": [{"fp": r[0], "N": r[1], "frac": round(r[4] / peak, 4)} for r in rows if r[4] / peak < 0.8],
        # every (precision, N) of the sweep: fraction of the measured HBM peak, median launch of the timed steps
        "per_size": {("f32" if fp == 4 else "f64"): {str(r[1]): round(r[4] / peak, 3) for r in rows if r[0] == fp}
                     for fp in (4, 8)},
    }
    if args.per_size and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.per_size)), exist_ok=True)
        with open(args.per_size, "w") as f:
            f.write("fp,N,K,time_us,GBs,frac_of_peak,GFLOPs,kernel,mean_time_us\n")
            for r in rows:
                f.write("%d,%d,%d,%.2f,%.1f,%.4f,%.1f,%s,%.2f\n" % (r[0], r[1], r[2], r[3] * 1e6, r[4], r[4] / peak, r[5], r[6], r[7] * 1e6))

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    e2e = None
    try:
        e2e = measure_e2e(args, torch, dist, plans, world, dev, barrier, total_flops)
    except Exception as ex:  # keep the device-resident numbers even if the host path fails
        e2e = {"value": None, "unit": "GFLOP/s", "error": str(ex)[:300]}

    sharded = None
    if not args.no_extra:
        try:
            sharded = sharded_configs(torch, dist, pkg, world, rank, dev, stream, peak)
        except Exception as ex:
            sharded = {"error": str(ex)[:300]}
            if world > 1:
                raise  # a rank that left the sequence would hang the others in the next barrier

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            try:
                cpu = run_pocketfft_sample(2, 1, cores)
            except Exception as ex:
                cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "unavailable", "sample": str(ex)[:200]}
            try:
                # the reference's generated kernels under the host emulator: parity-grade, kept beside it
                gf, cgbs, dt, kind, sample = run_reference_sample(1, 1, cores)
                cpu["reference_emulated"] = {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": kind,
                                             "sample": sample, "gbs": cgbs}
            except Exception as ex:  # the checker is optional for the measurement itself
                cpu["reference_emulated"] = {"value": None, "kind": "unavailable", "sample": str(ex)[:200]}
        other = None
        if world == 1 and not args.no_extra:
            try:
                other = other_configs(peak)
            except Exception as ex:
                other = {"error": str(ex)[:300]}
        line = {
            "metric": "c2c GFLOP/s (5N*log2N), 1d double-batched sweep N=2..512",
            "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64", "data": "synthetic",
            "config": workload_config(world),
            "gbs": gbs, "gbs_frac_of_peak": gbs / world / peak,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches_per_step * steps, "clocks": clocks,
            "other_configs": other,
            "sharded_configs": sharded,
        }
        emit(json.dumps(line))
    for _, _, _, p in plans:
        p.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
