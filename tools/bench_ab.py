"""A/B of planner candidates inside the real sweep: runs bench.py once per candidate set with
BBFFT_CUDA_WISDOM_OVERRIDE and prints, per size, the fraction of the HBM peak every set reached.
Usage: python tools/bench_ab.py [candidates.json]   ({"fp:N": [override, ...]}; writes gpurun_out/bench_ab.json)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load_candidates(path):
    """{"fp:N": [tune, ...]} from a JSON file; the entry currently in csrc/wisdom.inc is measured beside them."""
    import re
    wis = open(os.path.join(ROOT, "double-batched-fft-library_b200", "csrc", "wisdom.inc")).read()
    cand = {}
    for key, lst in json.load(open(path)).items():
        fp, n = (int(x) for x in key.split(":"))
        m = re.search(r'\{%d, %d, "([^"]*)"\}' % (fp, n), wis)
        cur = [m.group(1)] if m else []
        cand[(fp, n)] = cur + [t for t in lst if t not in cur]
    return cand


def main():
    CAND = load_candidates(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tools", "bench_ab_round2.json"))
    nsets = max(len(v) for v in CAND.values())
    res = {}
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    for rnd in range(2):  # two rounds: every set is measured twice, in alternating order
        for s in (range(nsets) if rnd == 0 else reversed(range(nsets))):
            ov = ";".join("%d:%d:%s" % (fp, n, c[min(s, len(c) - 1)]) for (fp, n), c in CAND.items())
            # every kernel of the sweep NVRTC-built, one per translation unit: candidates and incumbents in the
            # same build flavor (aot.py: NVRTC_BUILT must list the sizes whose winner is adopted from here)
            env = dict(os.environ, BBFFT_CUDA_WISDOM_OVERRIDE=ov, BBFFT_CUDA_JIT_LINEINFO="0", BBFFT_CUDA_NO_BUILTIN="1",
                       BBFFT_CUDA_KERNEL_CACHE=os.path.join(ROOT, "kcache"))
            path = os.path.join(out_dir, "bench_ab_%d_%d.csv" % (rnd, s))
            subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3", "--no-extra",
                            "--e2e-steps", "0", "--no-cpu-baseline", "--per-size", path], env=env, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=600)
            for r in csv.DictReader(open(path)):
                key = (int(r["fp"]), int(r["N"]))
                if key in CAND and s < len(CAND[key]):
                    res.setdefault("%d:%d" % key, {}).setdefault(CAND[key][s], []).append(float(r["frac_of_peak"]))
    json.dump(res, open(os.path.join(out_dir, "bench_ab.json"), "w"), indent=1)
    for k, v in res.items():
        best = max(v, key=lambda c: sum(v[c]) / len(v[c]))
        print(k, "  ".join("%s=%s" % (c, ["%.3f" % x for x in v[c]]) for c in v), " -> ", best)


if __name__ == "__main__":
    main()
