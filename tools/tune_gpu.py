"""Auto-tuner for the 1d c2c kernels (M=16 sweep family): times planner overrides on the GPU and
writes the best per (fp, N) as JSON "wisdom".  NVRTC compiles run on a thread pool.
Usage: python tools/tune_gpu.py --fp 4,8 --sizes 343,512 --out gpurun_out/wisdom.json [--minN 33]
"""
import argparse
import importlib
import itertools
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("double-batched-fft-library_b200")
aot = importlib.import_module("double-batched-fft-library_b200.aot")


def factorizations(n, max_r, max_l):
    out = []

    def rec(rem, lo, cur):
        if rem == 1:
            if cur:
                out.append(list(cur))
            return
        if len(cur) >= max_l:
            return
        for r in range(lo, min(rem, max_r) + 1):
            if rem % r == 0:
                cur.append(r)
                rec(rem // r, r, cur)
                cur.pop()
    rec(n, 2, [])
    return out


def candidates(n, fp, M):
    full = 128 // (2 * fp)
    cands = [""]
    single_max = 64 if fp == 4 else 32
    facs = []
    if n <= single_max:
        facs.append([n])
    for L, max_r, top in ((2, 32 if fp == 4 else 25, 3), (3, 16, 3), (4, 8, 1 if n > 300 else 0)):
        f = [x for x in factorizations(n, max_r, L) if len(x) == L]
        f.sort(key=lambda x: (max(x), sum(x)))
        facs += f[:top]
    seen = set()
    for f in facs:
        if len(f) == 1:
            for bh in (4, 8, 16):
                cands.append("R=%d,T=1,BH=%d" % (f[0], bh))
            continue
        ts = set()
        # threads per transform: one sub-FFT per thread in the largest-radix stage (and half of
        # that), and one per thread in EVERY stage (n / smallest radix: fewest registers)
        for t in (n // max(f), -(-(n // max(f)) // 2), n // min(f)):
            regs = max(-(-(n // r) // t) * r for r in f)
            if regs <= (32 if fp == 4 else 25):
                ts.add(t)
        for t in ts:
            for ml in {full, max(2, full // 2)}:
                if ml * t > 1024:
                    continue
                for mb in (1, 2, 3, 4):
                    bhs = {max(1, 256 // (ml * t)), max(1, 128 // (ml * t))}
                    for bh in bhs:
                        thr = ml * t * bh
                        if thr > 1024 or thr * mb > 2048:
                            continue
                        if ml * bh * n * 2 * fp * mb > 220 * 1024:
                            continue
                        s = "R=%s,T=%d,ML=%d,BH=%d,MB=%d" % ("x".join(map(str, f)), t, ml, bh, mb)
                        if s not in seen:
                            seen.add(s)
                            cands.append(s)
    return cands


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fp", default="4,8")
    ap.add_argument("--sizes", default="")
    ap.add_argument("--minN", type=int, default=2)
    ap.add_argument("--M", type=int, default=16)
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--out", default="gpurun_out/wisdom.json")
    ap.add_argument("--threads", type=int, default=min(32, os.cpu_count() or 8))
    ap.add_argument("--from-csv", default="", help="per-size CSV of bench.py: tune only the (fp, N) below --below")
    ap.add_argument("--below", type=float, default=0.9)
    ap.add_argument("--batch", type=int, default=24, help="sizes compiled ahead and then timed back to back")
    args = ap.parse_args()
    todo = []
    if args.from_csv:
        import csv
        for r in csv.DictReader(open(args.from_csv)):
            if float(r["frac_of_peak"]) < args.below:
                todo.append((int(r["fp"]), int(r["N"])))
    else:
        sizes = [int(s) for s in args.sizes.split(",")] if args.sizes else [n for n in aot.smooth_sizes() if n >= args.minN]
        todo = [(int(f), n) for f in args.fp.split(",") for n in sizes]
    stream = torch.cuda.current_stream().cuda_stream
    M = args.M
    results = {}
    pool = ThreadPoolExecutor(args.threads)
    xbuf = {4: torch.rand(args.bytes // 4, dtype=torch.float32, device="cuda"),
            8: torch.rand(args.bytes // 8, dtype=torch.float64, device="cuda")}
    ybuf = {4: torch.empty_like(xbuf[4]), 8: torch.empty_like(xbuf[8])}
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)

    def timed(plan, x, y, n):
        e0.record()
        for _ in range(n):
            plan.execute(x, y)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / n

    # Sustained-clock tuning: all candidates of a batch of sizes are compiled first, then timed
    # back to back without host gaps, so the GPU sits at its power-capped clocks like it does in
    # the benchmark sweep (burst timings favour register-starved, spill-heavy variants that lose
    # once the SM clock drops).
    for b0 in range(0, len(todo), args.batch):
        batch = todo[b0:b0 + args.batch]
        jobs = []
        for fp, n in batch:
            K = max(1, args.bytes // (M * n * 2 * fp))
            cfg = pkg.make_config(1, [M, n, K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
            for tune in candidates(n, fp, M):
                jobs.append((fp, n, K, cfg, tune))

        def mk(job):
            fp, n, K, cfg, tune = job
            try:
                return job, pkg.Plan(cfg, stream=stream, tune=tune)
            except Exception:
                return job, None
        built = [(j, p) for j, p in pool.map(mk, jobs) if p is not None]
        # heat up
        if built:
            j, p = built[0]
            for _ in range(300):
                p.execute(xbuf[j[0]], ybuf[j[0]])
        per = {}
        for (fp, n, K, cfg, tune), plan in built:
            x, y = xbuf[fp], ybuf[fp]
            plan.execute(x, y)
            t = timed(plan, x, y, 6)
            per.setdefault((fp, n), []).append([t, tune, plan, K])
        for (fp, n), lst in per.items():
            lst.sort(key=lambda t: t[0])
            x, y = xbuf[fp], ybuf[fp]
            finals = []
            for t, tune, plan, K in lst[:4]:
                finals.append([timed(plan, x, y, 40), tune, plan, K])
            finals.sort(key=lambda t: t[0])
            lst = finals + lst[4:]
            K = lst[0][3]
            nbytes = 2.0 * M * n * K * 2 * fp
            default = [t for t in lst if t[1] == ""]
            best = lst[0]
            results["%d,%d" % (fp, n)] = {"best": best[1], "gbs": nbytes / best[0] * 1e-6,
                                          "default_gbs": nbytes / default[0][0] * 1e-6 if default else None,
                                          "top": [(t[1], round(nbytes / t[0] * 1e-6)) for t in lst[:5]],
                                          "n_cands": len(lst), "mode": "sustained"}
            print(fp, n, "best %s %.0f GB/s (default %.0f) of %d cands; top: %s" % (
                best[1], nbytes / best[0] * 1e-6, nbytes / default[0][0] * 1e-6 if default else -1, len(lst),
                results["%d,%d" % (fp, n)]["top"][1:4]), flush=True)
        for _, plan in built:
            plan.close()
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
