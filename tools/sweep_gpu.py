"""GPU sweep over the C2 configuration family (1d c2c, M=16, seven-smooth N in [2,512]):
parity against numpy (small K) + bandwidth (K sized to ~1 GiB by default), CSV to stdout/file.
Usage: python tools/sweep_gpu.py [--fp 4,8] [--sizes 2,3,..] [--bytes 1073741824] [--tune "..."] [--out file.csv]
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("double-batched-fft-library_b200")


def smooth_sizes(lo=2, hi=512, primes=(2, 3, 5, 7)):
    out = []
    for n in range(lo, hi + 1):
        m = n
        for p in primes:
            while m % p == 0:
                m //= p
        if m == 1:
            out.append(n)
    return out


def time_plan(plan, x, y, reps=10, inner=3):
    # x, y are larger than L2 for the big configs; `inner` back-to-back launches per event pair
    for _ in range(3):
        plan.execute(x, y)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            plan.execute(x, y)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fp", default="4,8")
    ap.add_argument("--sizes", default="")
    ap.add_argument("--M", type=int, default=16)
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--tune", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--check", type=int, default=1)
    args = ap.parse_args()
    sizes = [int(s) for s in args.sizes.split(",")] if args.sizes else smooth_sizes()
    stream = torch.cuda.current_stream().cuda_stream
    rows = []
    hdr = "fp,M,N,K,radix,T,ML,BH,err,time_us,GBs,GFLOPs,kernel"
    print(hdr, flush=True)
    for fp in [int(f) for f in args.fp.split(",")]:
        dt = torch.complex64 if fp == 4 else torch.complex128
        esz = 2 * fp
        for N in sizes:
            M = args.M
            K = max(1, args.bytes // (M * N * esz))
            cfg = pkg.make_config(1, [M, N, K], fp, pkg.FORWARD, pkg.C2C, inplace=False)
            try:
                plan = pkg.Plan(cfg, stream=stream, tune=args.tune)
            except Exception as ex:
                print("# FAIL plan", fp, N, str(ex)[:200], flush=True)
                continue
            d = pkg.describe(cfg, args.tune)
            x = torch.randn(K, N, M, 2, dtype=torch.float32 if fp == 4 else torch.float64, device="cuda")
            x = torch.view_as_complex(x)
            y = torch.empty_like(x)
            err = -1.0
            if args.check:
                plan.execute(x, y)
                torch.cuda.synchronize()
                kc = min(K, 64)
                ref = torch.fft.fft(x[:kc].to(torch.complex128), dim=1)
                err = float((y[:kc].to(torch.complex128) - ref).norm() / ref.norm())
            t = time_plan(plan, x, y)
            nbytes = 2.0 * M * N * K * esz
            flops = 5.0 * N * np.log2(N) * M * K
            row = "%d,%d,%d,%d,%s,%d,%d,%d,%.2e,%.1f,%.1f,%.1f,%s" % (
                fp, M, N, K, "x".join(map(str, d["radix"])), d["threads_per_transform"], d["batch_lanes"],
                d["batch_high"], err, t * 1e6, nbytes / t * 1e-9, flops / t * 1e-9, d["identifier"])
            print(row, flush=True)
            rows.append(row)
            plan.close()
            del x, y
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(hdr + "\n" + "\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
