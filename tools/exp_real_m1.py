"""Experiment: planner overrides for the M=1 configurations (C1: c2c N=64; C3: r2c/c2r N=256)."""
import importlib, itertools, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("double-batched-fft-library_b200")
stream = torch.cuda.current_stream().cuda_stream


def timeit(plan, a, b, reps=7):
    for _ in range(2):
        plan.execute(a, b)
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); plan.execute(a, b); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def run(ttype, fp, N, K, tunes, inplace=False):
    nh = N // 2 + 1
    d = pkg.BACKWARD if ttype == pkg.C2R else pkg.FORWARD
    cfg = pkg.make_config(1, [1, N, K], fp, d, ttype, inplace=inplace)
    rdt = torch.float32 if fp == 4 else torch.float64
    if ttype == pkg.C2C:
        a = torch.rand(K * N * 2, dtype=rdt, device="cuda"); b = torch.empty_like(a)
        nbytes = 2.0 * N * 2 * fp * K
    elif ttype == pkg.R2C:
        a = torch.rand(K * N, dtype=rdt, device="cuda"); b = torch.empty(K * nh * 2, dtype=rdt, device="cuda")
        nbytes = float(N * fp + nh * 2 * fp) * K
    else:
        a = torch.rand(K * nh * 2, dtype=rdt, device="cuda"); b = torch.empty(K * N, dtype=rdt, device="cuda")
        nbytes = float(N * fp + nh * 2 * fp) * K
    res = []
    for t in tunes:
        try:
            plan = pkg.Plan(cfg, stream=stream, tune=t)
        except Exception as ex:
            print("  FAIL", t, str(ex)[:100]); continue
        dt = timeit(plan, a, b)
        res.append((nbytes / dt * 1e-9, t, plan.kernel_names[0]))
        plan.close()
    res.sort(reverse=True)
    for r in res[:8]:
        print("  %7.0f GB/s  %-40s %s" % r)
    return res


if __name__ == "__main__":
    K = 1 << 20
    print("c2r N=256 M=1")
    tunes = [""] + ["R=%s,T=%d,BH=%d,LD=%d,ST=%d,MB=%d" % (r, t, bh, ld, st, mb)
                    for r, t in (("8x16", 8), ("8x16", 16), ("16x8", 8), ("4x4x8", 16), ("4x32", 4), ("2x8x8", 16), ("8x16", 4))
                    for bh in (32 // max(1, t // 8), 16 // max(1, t // 8))
                    for ld in (0, 1) for st in (0, 1) for mb in (2, 4)]
    run(pkg.C2R, 4, 256, K, tunes)
    print("r2c N=256 M=1")
    run(pkg.R2C, 4, 256, K, tunes)
    print("c2c N=64 M=1 fp32")
    tunes = [""] + ["R=%s,T=%d,BH=%d,LD=%d,ST=%d,MB=%d" % (r, t, bh, ld, st, mb)
                    for r, t in (("8x8", 8), ("4x16", 4), ("16x4", 4), ("4x4x4", 16), ("8x8", 4))
                    for bh in (32, 16, 64) for ld in (0, 1) for st in (1,) for mb in (2, 4)]
    tunes += ["R=64,T=1,KL=1,BH=%d,MB=%d" % (bh, mb) for bh in (1, 2, 4) for mb in (1, 2, 3)]
    run(pkg.C2C, 4, 64, (1 << 30) // (64 * 8), tunes)
