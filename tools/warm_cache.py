"""Fill the persistent kernel cache (BBFFT_CUDA_KERNEL_CACHE) on a CPU-only box for a list of
(descriptor, overrides) pairs, so that a GPU run loads cubins instead of calling NVRTC.
Usage: python tools/warm_cache.py --tile-cases            (the candidates of tools/tune_tile.py)
       python tools/warm_cache.py "scfo16.64*100" "dcfo64x64*8:TH=512,MB=2" ..."""
import argparse
import ast
import importlib
import os
import re
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("items", nargs="*")
    ap.add_argument("--tile-cases", action="store_true")
    ap.add_argument("--real-sweep", action="store_true")
    ap.add_argument("--with-wisdom", action="store_true", help="plan like the library does by default (bench / sweep kernels)")
    ap.add_argument("--cases", action="append", default=[], help="JSON {descriptor: [overrides]} (tools/tune_list.py)")
    ap.add_argument("--dir", default=os.path.join(ROOT, "kcache"))
    args = ap.parse_args()
    os.makedirs(args.dir, exist_ok=True)
    os.environ["BBFFT_CUDA_KERNEL_CACHE"] = args.dir
    os.environ["BBFFT_CUDA_JIT_LINEINFO"] = "0"
    # candidates are planned without the measured tables on both sides (tools/tune_list.py sets the same):
    # a wisdom entry merged into an override string here but not on the GPU box would change the kernel
    # and with it the cache key (that cost the config-3 search of round 1 its GPU time)
    if not args.with_wisdom:
        os.environ["BBFFT_CUDA_NO_WISDOM"] = "1"
    pkg = importlib.import_module("double-batched-fft-library_b200")
    jobs = [tuple(i.split(":", 1)) if ":" in i else (i, "") for i in args.items]
    if args.real_sweep:
        # the configurations of tools/bench_configs.py --real-sweep
        aot = importlib.import_module("double-batched-fft-library_b200.aot")
        sizes = [n for n in aot.smooth_sizes() if n in (2, 3, 4, 7, 8, 15, 16, 27, 32, 49, 64, 100, 105, 128, 135, 200, 243, 256,
                                                         315, 343, 384, 400, 441, 480, 500, 512)]
        for fp in (4, 8):
            for n in sizes:
                k = max(2, (1 << 30) // (16 * n * fp)) // 2 * 2
                for t in ("f", "b"):
                    jobs.append(("%sr%so16.%d*%d" % ("s" if fp == 4 else "d", t, n, k), ""))
    for path in args.cases:
        import json
        for d, ts in json.load(open(path)).items():
            jobs += [(d, t) for t in ts]
    if args.tile_cases:
        src = open(os.path.join(ROOT, "tools", "tune_tile.py")).read()
        cases = ast.literal_eval(re.search(r"CASES = (\{.*?\n\})", src, re.S).group(1))
        jobs += [(d, t) for d, ts in cases.items() for t in ts]

    def one(job):
        desc, tune = job
        try:
            d = pkg.describe(pkg.parse_descriptor(desc), tune)
        except Exception:
            return None  # the planner rejects the override
        pkg.compile_to_cubin(d["source"])
        return d["identifier"]

    with ThreadPoolExecutor(os.cpu_count() or 4) as pool:
        names = list(pool.map(one, jobs))
    print("%d kernels in %s" % (len(set(n for n in names if n)), args.dir))


if __name__ == "__main__":
    main()
