"""Auto-tuner for the 1d kernels (M=16 sweep family, c2c / r2c / c2r): times planner overrides on
the GPU and writes the best per (type, fp, N) as JSON "wisdom" (tools/make_wisdom.py turns it into
csrc/wisdom*.inc).  Candidates are either enumerated here and NVRTC-compiled on a thread pool, or
read from tools/tune_prepare.py's lists, whose kernels are already in the persistent kernel cache.
Usage: python tools/tune_gpu.py --type c2c --fp 4,8 --sizes 343,512 --out gpurun_out/wisdom.json
       BBFFT_CUDA_KERNEL_CACHE=tune_cache BBFFT_CUDA_JIT_LINEINFO=0 BBFFT_CUDA_NO_WISDOM=1 \
           python tools/tune_gpu.py --cands tune_cache/cands_c2c.json --out gpurun_out/w.json
"""
import argparse
import importlib
import itertools
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


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


def candidates(n, fp, M, ttype="c2c"):
    """Planner overrides to time for one (type, fp, N).  Real transforms run their stages on the
    complex length (N/2 for even N); their fused pre/post pass holds a sub-FFT AND its mirror in
    registers in the first (c2r) / last (r2c) stage, so that stage gets the smallest radix."""
    full = 128 // (2 * fp)
    cands = [""]
    single_max = 64 if fp == 4 else 32
    if ttype != "c2c" and n % 2 == 0:
        n = n // 2
    facs = []
    if n <= single_max:
        facs.append([n])
    real = ttype != "c2c"  # real sweeps are twice as many configurations: a leaner candidate set
    for L, max_r, top in ((2, 32 if fp == 4 else 25, 2 if real else 3), (3, 16, (2 if n > 100 else 1) if real else 3),
                          (4, 8, 1 if n > 300 and not real else 0)):
        f = [x for x in factorizations(n, max_r, L) if len(x) == L]
        f.sort(key=lambda x: (max(x), sum(x)))
        facs += f[:top]
    if ttype == "r2c":
        facs = [sorted(f, reverse=True) for f in facs]
    seen = set()
    for f in facs:
        if len(f) == 1:
            for bh in (4, 8, 16):
                cands.append("R=%d,T=1,BH=%d" % (f[0], bh))
            continue
        ts = set()
        # threads per transform: one sub-FFT per thread in the largest-radix stage (and half of
        # that), and one per thread in EVERY stage (n / smallest radix: fewest registers)
        tlist = [n // max(f), -(-(n // max(f)) // 2), n // min(f)]
        if ttype != "c2c":
            # one mirrored unit (two sub-FFTs) per thread in the fused pre/post stage
            rm = f[-1] if ttype == "r2c" else f[0]
            tlist = [n // max(f), (n // rm) // 2 + 1, n // min(f)]
        for t in tlist:
            regs = max(-(-(n // r) // t) * r for r in f)
            if regs <= (32 if fp == 4 else 25):
                ts.add(t)
        for t in ts:
            # real transforms keep full lanes: a CTA that owns every m of a k slice is what makes
            # the in-place variant legal (planner: inplace_unsupported)
            for ml in ({full, max(2, full // 2)} if ttype == "c2c" else {full}):
                if ml * t > 1024:
                    continue
                for mb in (1, 2, 3, 4):
                    bhs = {max(1, 256 // (ml * t)), max(1, 128 // (ml * t))}
                    if real:
                        if mb == 1 and ml * t < 512:
                            continue
                        if ml * t > 64:
                            bhs = {max(1, 256 // (ml * t))}
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
                            # a wide mirrored stage is register-hungry: also time the pre/post
                            # pass as a separate trip through shared memory
                            if ttype != "c2c" and mb >= 2 and (f[-1] if ttype == "r2c" else f[0]) > 8:
                                cands.append(s + ",RF=0")
    return cands


def make_cfg(pkg, ttype, fp, n, M, nbytes):
    """The sweep configuration of one size: input tensor of `nbytes` (benchmark/test.hpp:18-30)."""
    if ttype == "c2c":
        K = max(1, nbytes // (M * n * 2 * fp))
        return pkg.make_config(1, [M, n, K], fp, pkg.FORWARD, pkg.C2C, inplace=False), K
    slack = nbytes // 16 + (1 << 20)
    K = max(2, nbytes // (M * n * fp)) // 2 * 2
    while K > 2 and M * (n // 2 + 1) * K * 2 * fp > nbytes + slack:
        K -= 2
    return pkg.make_config(1, [M, n, K], fp, pkg.FORWARD if ttype == "r2c" else pkg.BACKWARD,
                           pkg.R2C if ttype == "r2c" else pkg.C2R, inplace=False), K


def algorithmic_bytes(ttype, fp, n, M, K):
    if ttype == "c2c":
        return 2.0 * M * n * K * 2 * fp
    return float(n * fp + (n // 2 + 1) * 2 * fp) * M * K


def main():
    import torch
    pkg = importlib.import_module("double-batched-fft-library_b200")
    aot = importlib.import_module("double-batched-fft-library_b200.aot")
    ap = argparse.ArgumentParser()
    ap.add_argument("--fp", default="4,8")
    ap.add_argument("--sizes", default="")
    ap.add_argument("--minN", type=int, default=2)
    ap.add_argument("--M", type=int, default=16)
    ap.add_argument("--type", default="c2c", help="comma list of c2c,r2c,c2r")
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--out", default="gpurun_out/wisdom.json")
    ap.add_argument("--threads", type=int, default=min(32, os.cpu_count() or 8))
    ap.add_argument("--from-csv", default="", help="per-size CSV of bench.py: tune only the (fp, N) below --below")
    ap.add_argument("--below", type=float, default=0.9)
    ap.add_argument("--cands", default="", help="candidate lists written by tools/tune_prepare.py (comma list of files)")
    ap.add_argument("--reps", type=int, default=7, help="interleaved timing passes over all candidates")
    ap.add_argument("--chunk", type=int, default=4000, help="candidates resident at once")
    ap.add_argument("--filler", type=int, default=0,
                    help="untimed launches of HBM-saturating sweep kernels after every candidate (power/clock regime of the sweep)")
    args = ap.parse_args()
    todo = {}  # (type, fp, n) -> [tune, ...]
    if args.cands:
        for path in args.cands.split(","):
            for key, tunes in json.load(open(path)).items():
                t, fp, n = key.split(",")
                have = todo.setdefault((t, int(fp), int(n)), [])
                have += [x for x in tunes if x not in have]
    else:
        pairs = []
        if args.from_csv:
            import csv
            for r in csv.DictReader(open(args.from_csv)):
                if float(r["frac_of_peak"]) < args.below:
                    pairs.append((int(r["fp"]), int(r["N"])))
        else:
            sizes = [int(x) for x in args.sizes.split(",")] if args.sizes else [n for n in aot.smooth_sizes() if n >= args.minN]
            pairs = [(int(f), n) for f in args.fp.split(",") for n in sizes]
        for t in args.type.split(","):
            for fp, n in pairs:
                todo[(t, fp, n)] = candidates(n, fp, args.M, t)
    stream = torch.cuda.current_stream().cuda_stream
    M = args.M
    pool = ThreadPoolExecutor(args.threads)
    slack = args.bytes // 16 + (1 << 20)  # the spectrum of a real transform is (N/2+1)/(N/2) larger
    xbuf = {4: torch.rand((args.bytes + slack) // 4, dtype=torch.float32, device="cuda"),
            8: torch.rand((args.bytes + slack) // 8, dtype=torch.float64, device="cuda")}
    ybuf = {4: torch.empty_like(xbuf[4]), 8: torch.empty_like(xbuf[8])}

    # Interleaved timing: one launch of every candidate in turn, `reps` passes, CUDA events per
    # launch.  The SM clock under the power cap is then set by the whole mix -- the regime of the
    # benchmark sweep, where each kernel runs once between others -- instead of by the candidate
    # itself (40 back-to-back launches of a frugal kernel climb to boost clocks and look 15-25 %
    # faster than they are in the sweep: profiles/r01_tuner4_sustained.json vs r01e_per_size.csv).
    jobs = []
    for (t, fp, n), tunes in sorted(todo.items()):
        cfg, K = make_cfg(pkg, t, fp, n, M, args.bytes)
        for tune in tunes:
            jobs.append((t, fp, n, K, cfg, tune))
    print("%d candidates, %d configurations" % (len(jobs), len(todo)), flush=True)
    # Filler: the benchmark sweep is dominated by kernels that saturate HBM; they set the board power
    # and with it the SM clock (~1610 MHz under the 1000 W cap) that the slower, compute-heavier
    # kernels then have to live with.  Timing weak sizes only among themselves lets the clock rise
    # and over-predicts them by ~10 % (profiles/r01f_tuner_interleaved.json vs r01g_per_size.csv).
    fillers = []
    if args.filler > 0:
        for fp, n in ((4, 64), (8, 64), (4, 16), (8, 128)):
            cfg, K = make_cfg(pkg, "c2c", fp, n, M, args.bytes)
            os.environ.pop("BBFFT_CUDA_NO_WISDOM", None)
            fillers.append((fp, pkg.Plan(cfg, stream=stream)))
        os.environ["BBFFT_CUDA_NO_WISDOM"] = "1"
    nfill = 0
    times = {}  # job index -> [ms, ...]
    t_start = time.time()
    for c0 in range(0, len(jobs), args.chunk):
        chunk = list(enumerate(jobs[c0:c0 + args.chunk], start=c0))

        def mk(item):
            idx, (t, fp, n, K, cfg, tune) = item
            try:
                return idx, pkg.Plan(cfg, stream=stream, tune=tune)
            except Exception:
                return idx, None
        built = [(i, p) for i, p in pool.map(mk, chunk) if p is not None]
        print("chunk %d: %d plans built after %.0f s" % (c0 // args.chunk, len(built), time.time() - t_start), flush=True)
        if not built:
            continue
        order = list(range(len(built)))
        # heat up with one untimed pass, then the timed passes in a shuffled order per pass
        import random
        rnd = random.Random(1234)
        for rep in range(args.reps + 1):
            evs = []
            rnd.shuffle(order)
            for o in order:
                idx, plan = built[o]
                fp = jobs[idx][1]
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                plan.execute(xbuf[fp], ybuf[fp])
                e1.record()
                evs.append((idx, e0, e1))
                for _ in range(args.filler):
                    ffp, fplan = fillers[nfill % len(fillers)]
                    fplan.execute(xbuf[ffp], ybuf[ffp])
                    nfill += 1
            torch.cuda.synchronize()
            if rep > 0:
                for idx, e0, e1 in evs:
                    times.setdefault(idx, []).append(e0.elapsed_time(e1))
        for _, plan in built:
            plan.close()
    results = {}
    per = {}
    for idx, ts in times.items():
        t, fp, n, K, cfg, tune = jobs[idx]
        ts.sort()
        per.setdefault((t, fp, n), []).append((ts[len(ts) // 2], tune, K))
    for (t, fp, n), lst in sorted(per.items()):
        lst.sort(key=lambda x: x[0])
        K = lst[0][2]
        nbytes = algorithmic_bytes(t, fp, n, M, K)
        default = [x for x in lst if x[1] == ""]
        best = lst[0]
        key = "%d,%d" % (fp, n) if t == "c2c" else "%s,%d,%d" % (t, fp, n)
        results[key] = {"best": best[1], "gbs": nbytes / best[0] * 1e-6,
                        "default_gbs": nbytes / default[0][0] * 1e-6 if default else None,
                        "top": [(x[1], round(nbytes / x[0] * 1e-6)) for x in lst[:5]],
                        "n_cands": len(lst), "mode": "interleaved"}
        print(t, fp, n, "best %s %.0f GB/s (default %.0f) of %d cands; top: %s" % (
            best[1], nbytes / best[0] * 1e-6, nbytes / default[0][0] * 1e-6 if default else -1, len(lst),
            results[key]["top"][1:4]), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)
    print("done in %.0f s" % (time.time() - t_start))


if __name__ == "__main__":
    main()
