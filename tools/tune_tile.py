"""Time planner overrides of the fused 2d tile kernel (bbk::fft2d_tile) on the GPU.
Usage: python tools/tune_tile.py [--out gpurun_out/tile_tune.json]"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("double-batched-fft-library_b200")

CASES = {
    # descriptor: candidate overrides ("" = planner default)
    "dcfo64x64*4096": ["", "RA=8x8,RB=8x8", "TH=128,MB=2", "TH=512,MB=2", "RA=2x32,RB=2x32", "RA=4x16,RB=8x8", "RA=8x8,RB=4x16",
                       "RA=4x16,RB=4x16,PADK=8", "RA=4x16,RB=4x16,PADK=32", "RA=4x16,RB=4x16,PADK=64", "RA=4x16,RB=4x16,MB=1"],
    "scfo128x128*8192": ["", "RA=4x32,RB=4x32", "RA=8x16,RB=4x32", "RA=4x32,RB=8x16", "RA=4x32,RB=4x32,TH=512,MB=1"],
    "scfo64x64*32768": ["", "RA=8x8,RB=8x8", "TH=128,MB=4", "TH=128,MB=6", "TH=256,MB=4", "TH=64,MB=6", "RA=2x32,RB=2x32",
                        "RA=2x32,RB=2x32,TH=128,MB=4"],
    "dcfo32x32*65536": ["", "RA=2x16,RB=2x16", "RA=2x16,RB=2x16,TH=64,MB=8"],
    "scfo32x32*131072": ["", "RA=2x16,RB=2x16", "RA=32,RB=32", "TH=128,MB=8"],
    "scfo256x64*8192": ["", "RA=8x32,RB=4x16", "RA=16x16,RB=8x8"],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/tile_tune.json")
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    stream = torch.cuda.current_stream().cuda_stream
    built = []
    for desc, tunes in CASES.items():
        cfg = pkg.parse_descriptor(desc)
        n = 1
        for d in range(cfg.dim + 2):
            n *= cfg.shape[d]
        for tune in tunes:
            try:
                built.append((desc, tune, cfg.fp, n, pkg.Plan(cfg, stream=stream, tune=tune)))
            except Exception as ex:
                print(desc, tune, "rejected:", str(ex)[:100], flush=True)
    nmax = max(b[3] * 2 * b[2] for b in built)
    x = torch.rand(nmax // 4, dtype=torch.float32, device="cuda")
    y = torch.empty_like(x)
    times = {}
    for rep in range(args.reps + 1):
        evs = []
        for i, (desc, tune, fp, n, plan) in enumerate(built):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.execute(x, y)
            e1.record()
            evs.append((i, e0, e1))
        torch.cuda.synchronize()
        if rep:
            for i, e0, e1 in evs:
                times.setdefault(i, []).append(e0.elapsed_time(e1))
    out = {}
    for i, (desc, tune, fp, n, plan) in enumerate(built):
        ts = sorted(times[i])
        t = ts[len(ts) // 2]
        gbs = 2.0 * n * 2 * fp / t * 1e-6
        out.setdefault(desc, []).append((tune, round(gbs), round(t * 1e3, 1), plan.kernel_names[0]))
        plan.close()
    for desc, lst in out.items():
        lst.sort(key=lambda r: -r[1])
        for tune, gbs, us, name in lst:
            print("%-22s %-36s %6d GB/s %8.1f us  %s" % (desc, tune or "(default)", gbs, us, name), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
