"""Run every measured planner entry (csrc/wisdom.inc, csrc/wisdom_real.inc) through the CPU emulator
against the oracle at M = 16 (the shape the entries were measured on) and a small K: a wrong table
entry would otherwise only show up on the GPU.  Test infrastructure.
Usage: python tools/check_wisdom_emu.py [--real-only] [--stride 1]"""
import argparse
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--real-only", action="store_true")
    ap.add_argument("--stride", type=int, default=1, help="check every stride-th entry")
    args = ap.parse_args()
    import numpy as np
    import emu
    from oracle import oracle
    from common import TOL, rel_l2
    pkg = emu.pkg
    csrc = os.path.join(ROOT, "double-batched-fft-library_b200", "csrc")
    entries = []
    if not args.real_only:
        for line in open(os.path.join(csrc, "wisdom.inc")):
            m = re.match(r'\{(\d+), (\d+), "([^"]*)"\}', line)
            if m:
                entries.append((0, int(m.group(1)), int(m.group(2)), m.group(3)))
    for line in open(os.path.join(csrc, "wisdom_real.inc")):
        m = re.match(r'\{(\d+), (\d+), (\d+), 0, "([^"]*)"\}', line)
        if m:
            entries.append((int(m.group(1)), int(m.group(2)), int(m.group(3)), m.group(4)))
    bad = 0
    M, K = 16, 2
    for i, (ttype, fp, N, tune) in enumerate(entries[::args.stride]):
        d = 1 if ttype == 2 else -1
        cfg = pkg.make_config(1, [M, N, K], fp, d, ttype, inplace=False)
        desc = pkg.describe(cfg)  # planned WITH wisdom: the entry must be what the planner uses
        r = re.search(r"R=([0-9x]+)", tune)
        if r and ("_r%s_" % r.group(1)) not in desc["identifier"]:
            print("entry not applied:", ttype, fp, N, tune, desc["identifier"][:80])
            bad += 1
        ocfg = oracle.make_config(1, [M, N, K], fp, d, ttype, inplace=False)
        rng = np.random.default_rng(i)
        rdt = np.float32 if fp == 4 else np.float64
        cdt = np.complex64 if fp == 4 else np.complex128
        nh = N // 2 + 1
        if ttype == 1:
            x = rng.standard_normal(M * N * K).astype(rdt)
            got, want = np.zeros(M * nh * K, cdt), np.zeros(M * nh * K, cdt)
        elif ttype == 2:
            x = (rng.standard_normal((K, nh, M)) + 1j * rng.standard_normal((K, nh, M))).astype(cdt)
            if N % 2 == 0:
                x[:, nh - 1, :] = x[:, nh - 1, :].real
            x = x.reshape(-1)
            got, want = np.zeros(M * N * K, rdt), np.zeros(M * N * K, rdt)
        else:
            x = (rng.standard_normal(M * N * K) + 1j * rng.standard_normal(M * N * K)).astype(cdt)
            got, want = np.zeros_like(x), np.zeros_like(x)
        emu.run(cfg, x, got)
        oracle.dft(ocfg, x, want)
        err = rel_l2(got, want)
        ok = err < TOL[fp] * 0.5
        bad += 0 if ok else 1
        if not ok or i % 40 == 0:
            print("%s type=%d fp=%d N=%d %s err=%.2e" % ("ok " if ok else "BAD", ttype, fp, N, tune, err), flush=True)
    print("checked %d entries: %d problems" % (len(entries[::args.stride]), bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
