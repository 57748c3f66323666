"""Re-key pre-compiled tuner candidates after the device header changed.

The persistent kernel cache (csrc/runtime.cpp) keys a cubin by hash(device header, stub source,
arch, options).  tools/tune_prepare.py compiles thousands of candidates; when bbfft_kernels.cuh is
edited afterwards in a way that does not touch the 1d kernels (e.g. a new nd kernel was added) the
cubins are still valid but their keys are not.  This script recomputes both keys for every kept
candidate (old header text from a git revision, new header from the working tree) and copies the
cubin to its new name in a fresh directory that then travels to the GPU box.

Usage: python tools/tune_rekey.py --old-rev 7fe6eca --src tune_cache --dst tune_cache_gpu cands_*.json
"""
import argparse
import importlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
MASK = (1 << 64) - 1
HDR = "double-batched-fft-library_b200/csrc/kernels/bbfft_kernels.cuh"


def fnv(h, data):
    for b in data:
        h = ((h ^ b) * 0x100000001b3) & MASK
    return h


def key(hdr_state, source, arch=b"sm_100a", nolineinfo=True):
    h1 = fnv(hdr_state, source)
    h2 = fnv(0x84222325cbf29ce4, source)
    h2 = fnv(h2, arch)
    if nolineinfo:
        h2 = fnv(h2, b"nolineinfo")
    return "%016x%016x.cubin" % (h1, h2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cands", nargs="+")
    ap.add_argument("--old-rev", required=True)
    ap.add_argument("--src", default=os.path.join(ROOT, "tune_cache"))
    ap.add_argument("--dst", default=os.path.join(ROOT, "tune_cache_gpu"))
    args = ap.parse_args()
    os.environ["BBFFT_CUDA_NO_WISDOM"] = "1"
    pkg = importlib.import_module("double-batched-fft-library_b200")
    import tune_gpu
    old_hdr = subprocess.run(["git", "show", "%s:%s" % (args.old_rev, HDR)], cwd=ROOT, capture_output=True, check=True).stdout
    new_hdr = open(os.path.join(ROOT, HDR), "rb").read()
    # the header is embedded as a C string: same bytes as the file
    assert new_hdr == pkg.kernel_header().encode(), "rebuild the library first (embedded header is stale)"
    st_old = fnv(0xcbf29ce484222325, old_hdr)
    st_new = fnv(0xcbf29ce484222325, new_hdr)
    os.makedirs(args.dst, exist_ok=True)
    n = miss = 0
    for path in args.cands:
        cands = json.load(open(path))
        for k, tunes in cands.items():
            t, fp, size = k.split(",")
            cfg, _ = tune_gpu.make_cfg(pkg, t, int(fp), int(size), 16, 1 << 30)
            for tune in tunes:
                src = pkg.describe(cfg, tune)["source"].encode()
                new = key(st_new, src)
                for cand in (new, key(st_old, src)):
                    p = os.path.join(args.src, cand)
                    if os.path.exists(p):
                        shutil.copyfile(p, os.path.join(args.dst, new))
                        n += 1
                        break
                else:
                    miss += 1
        shutil.copyfile(path, os.path.join(args.dst, os.path.basename(path)))
    print("copied %d cubins to %s (%d candidates without a cubin will be JIT-compiled)" % (n, args.dst, miss))


if __name__ == "__main__":
    main()
