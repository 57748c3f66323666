"""Pre-compile the auto-tuner's candidates on a CPU-only box.

For every (type, fp, N) of the M=16 sweep family this enumerates the planner overrides
tools/tune_gpu.py would time, compiles each kernel with NVRTC into the persistent kernel cache
(BBFFT_CUDA_KERNEL_CACHE, csrc/runtime.cpp) and drops the candidates whose cubin spills heavily.
The cache directory travels to the GPU box with the repo snapshot, so the GPU-side tuner
(tools/tune_gpu.py --cands ...) spends its time measuring, not compiling.

Usage: python tools/tune_prepare.py --type c2c,r2c,c2r --fp 4,8 --minN 33 --out tune_cache
"""
import argparse
import importlib
import json
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--type", default="c2c,r2c,c2r")
    ap.add_argument("--fp", default="4,8")
    ap.add_argument("--sizes", default="")
    ap.add_argument("--minN", type=int, default=33)
    ap.add_argument("--M", type=int, default=16)
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--from-csv", default="", help="per-size CSV of bench.py: only the c2c (fp, N) below --below")
    ap.add_argument("--below", type=float, default=0.9)
    ap.add_argument("--tag", default="")
    ap.add_argument("--only-wisdom", action="store_true", help="compile only the current wisdom entries of the sizes")
    ap.add_argument("--max-stack", type=int, default=64, help="drop candidates that spill more than this many bytes")
    ap.add_argument("--out", default=os.path.join(ROOT, "tune_cache"))
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    os.environ["BBFFT_CUDA_KERNEL_CACHE"] = args.out
    os.environ["BBFFT_CUDA_JIT_LINEINFO"] = "0"
    os.environ["BBFFT_CUDA_NO_WISDOM"] = "1"
    pkg = importlib.import_module("double-batched-fft-library_b200")
    aot = importlib.import_module("double-batched-fft-library_b200.aot")
    import tune_gpu

    sizes = [int(s) for s in args.sizes.split(",")] if args.sizes else [n for n in aot.smooth_sizes() if n >= args.minN]
    wisdom = {}
    import re as _re
    for line in open(os.path.join(ROOT, "double-batched-fft-library_b200", "csrc", "wisdom.inc")):
        m = _re.match(r'\{(\d+), (\d+), "([^"]*)"\}', line)
        if m:
            wisdom[(int(m.group(1)), int(m.group(2)))] = m.group(3)
    jobs = []
    pairs = [(int(f), n) for f in args.fp.split(",") for n in sizes]
    if args.from_csv:
        import csv
        pairs = [(int(r["fp"]), int(r["N"])) for r in csv.DictReader(open(args.from_csv)) if float(r["frac_of_peak"]) < args.below]
    for ttype in args.type.split(","):
        for fp, n in pairs:
            if True:
                cfg, K = tune_gpu.make_cfg(pkg, ttype, fp, n, args.M, args.bytes)
                tunes = [] if args.only_wisdom else tune_gpu.candidates(n, fp, args.M, ttype)
                # the entry the library uses today competes as well, so a re-tune can never regress
                if ttype == "c2c" and (fp, n) in wisdom and wisdom[(fp, n)] not in tunes:
                    tunes.append(wisdom[(fp, n)])
                for tune in tunes:
                    jobs.append((ttype, fp, n, cfg, tune))
    print("%d candidates to compile" % len(jobs), flush=True)

    def one(job):
        ttype, fp, n, cfg, tune = job
        try:
            d = pkg.describe(cfg, tune)
            cubin = pkg.compile_to_cubin(d["source"])
        except Exception as ex:  # planner rejects the override (CTA too large, smem, ...)
            return job, None
        with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
            f.write(cubin)
            f.flush()
            out = subprocess.run(["cuobjdump", "-res-usage", f.name], capture_output=True, text=True).stdout
        m = re.search(r"REG:(\d+) STACK:(\d+)", out)
        return job, (d["identifier"], int(m.group(1)), int(m.group(2)))

    cands = {}
    seen = set()
    done = 0
    with ThreadPoolExecutor(args.threads) as pool:
        for job, res in pool.map(one, jobs):
            done += 1
            if done % 500 == 0:
                print("  %d / %d" % (done, len(jobs)), flush=True)
            if res is None:
                continue
            ttype, fp, n, cfg, tune = job
            ident, reg, stack = res
            # different override strings can plan the same kernel; the default ("") always stays
            if tune and (stack > args.max_stack or ident in seen):
                continue
            seen.add(ident)
            cands.setdefault("%s,%d,%d" % (ttype, fp, n), []).append(tune)
    path = os.path.join(args.out, "cands_%s%s.json" % (args.type.replace(",", "_"), args.tag))
    with open(path, "w") as f:
        json.dump(cands, f)
    print("kept %d candidates for %d configurations -> %s" % (sum(len(v) for v in cands.values()), len(cands), path))


if __name__ == "__main__":
    main()
