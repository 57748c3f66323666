"""Time planner overrides for an explicit list of configurations (1d or fused-2d descriptors) on the
GPU: every candidate once per pass, several passes, CUDA events per launch (tools/tune_gpu.py's
interleaved regime), HBM-saturating filler launches in between.
Usage: python tools/tune_list.py --cases tools/cases_c3.json --out gpurun_out/c3_tune.json [--inner 1]
The cases file maps a descriptor ("srfo256*1048576") to a list of override strings ("" = default)."""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# same planning environment as tools/warm_cache.py, so that pre-compiled candidates are cache hits
os.environ["BBFFT_CUDA_NO_WISDOM"] = "1"
os.environ.setdefault("BBFFT_CUDA_JIT_LINEINFO", "0")
# candidates that spill under their cap are timed as they are (no probe, no second compile on the GPU box)
os.environ.setdefault("BBFFT_CUDA_KEEP_REGCAP", "1")
if os.path.isdir(os.path.join(ROOT, "kcache")):
    os.environ.setdefault("BBFFT_CUDA_KERNEL_CACHE", os.path.join(ROOT, "kcache"))
pkg = importlib.import_module("double-batched-fft-library_b200")


def tensor_bytes(cfg, inplace):
    n = 1
    for d in range(cfg.dim + 2):
        n *= cfg.shape[d]
    real_in = cfg.type == pkg.R2C
    real_out = cfg.type == pkg.C2R
    n1 = cfg.shape[1]
    other = n // n1
    spec = (n1 // 2 + 1) if cfg.type != pkg.C2C else n1
    cbytes = other * spec * 2 * cfg.fp
    rbytes = other * (2 * spec if inplace else n1) * cfg.fp
    ib = rbytes if real_in else cbytes
    ob = rbytes if real_out else cbytes
    alg = (other * n1 * cfg.fp + cbytes) if cfg.type != pkg.C2C else 2 * cbytes
    return ib, ob, alg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", required=True)
    ap.add_argument("--out", default="gpurun_out/tune_list.json")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--inner", type=int, default=1, help="launches per timed sample (L2-resident configurations: 20)")
    ap.add_argument("--filler", type=int, default=1)
    ap.add_argument("--graph", action="store_true",
                    help="time `inner` executes replayed as one CUDA graph (launch-bound shapes: what the GPU needs per "
                         "launch, without the Python call overhead)")
    args = ap.parse_args()
    cases = json.load(open(args.cases))
    stream = torch.cuda.current_stream().cuda_stream
    built = []
    nmax = 1 << 30
    for desc, tunes in cases.items():
        cfg = pkg.parse_descriptor(desc)
        inplace = "i" in desc[:4]
        ib, ob, alg = tensor_bytes(cfg, inplace)
        nmax = max(nmax, ib, ob)
        for tune in tunes:
            try:
                built.append((desc, tune, inplace, alg, pkg.Plan(cfg, stream=stream, tune=tune)))
            except Exception as ex:
                print(desc, tune, "rejected:", str(ex)[:100], flush=True)
    x = torch.rand(nmax // 4 + 1024, dtype=torch.float32, device="cuda")
    y = torch.empty_like(x)
    fillers = []
    if args.filler:
        os.environ.pop("BBFFT_CUDA_NO_WISDOM", None)  # the sweep's own kernels (built-in bundle)
        for d in ("scfo16.64*131072", "dcfo16.64*65536"):
            fillers.append(pkg.Plan(pkg.parse_descriptor(d), stream=stream))
        os.environ["BBFFT_CUDA_NO_WISDOM"] = "1"
    graphs = {}
    if args.graph:
        side = torch.cuda.Stream()
        for i, (desc, tune, inplace, alg, plan) in enumerate(built):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(args.inner):
                    if inplace:
                        plan.execute(x, stream=torch.cuda.current_stream().cuda_stream)
                    else:
                        plan.execute(x, y, stream=torch.cuda.current_stream().cuda_stream)
            graphs[i] = g
    times = {}
    nf = 0
    import random
    rnd = random.Random(7)
    order = list(range(len(built)))
    for rep in range(args.reps + 1):
        rnd.shuffle(order)
        evs = []
        for i in order:
            desc, tune, inplace, alg, plan = built[i]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if args.graph:
                graphs[i].replay()
            else:
                for _ in range(args.inner):
                    if inplace:
                        plan.execute(x)
                    else:
                        plan.execute(x, y)
            e1.record()
            evs.append((i, e0, e1))
            for _ in range(args.filler):
                fillers[nf % len(fillers)].execute(x, y)
                nf += 1
        torch.cuda.synchronize()
        if rep:
            for i, e0, e1 in evs:
                times.setdefault(i, []).append(e0.elapsed_time(e1) / args.inner)
    out = {}
    for i, (desc, tune, inplace, alg, plan) in enumerate(built):
        ts = sorted(times[i])
        t = ts[len(ts) // 2]
        out.setdefault(desc, []).append((tune, round(alg / t * 1e-6), round(t * 1e3, 2), plan.kernel_names[0]))
        plan.close()
    for desc, lst in out.items():
        lst.sort(key=lambda r: -r[1])
        for tune, gbs, us, name in lst[:12]:
            print("%-22s %-44s %6d GB/s %9.2f us  %s" % (desc, tune or "(default)", gbs, us, name[:60]), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
