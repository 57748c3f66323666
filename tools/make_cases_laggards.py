"""Cases file (tools/tune_list.py format) for the sizes that sit below 0.8 of the HBM peak in the sweeps
of round 2 (profiles/r02e_tile_pdl.txt, r02f_cluster.txt): the tuner's own candidates (tools/tune_gpu.py)
plus a wider net -- quarter lanes, more factorizations per stage count, 2-stage splits with a radix up
to 36 in fp32, every thread count that deals the sub-FFTs evenly.
Usage: python tools/make_cases_laggards.py > tools/cases_laggards.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tune_gpu  # noqa: E402

ROUND2 = [(8, 100), (8, 243), (8, 405), (8, 441), (8, 250), (8, 294), (8, 343), (8, 448), (4, 441), (4, 450), (4, 300), (4, 108),
          (8, 245), (8, 315), (4, 500), (8, 504), (8, 420), (4, 486), (8, 512), (4, 375), (4, 245), (8, 175), (4, 405), (4, 315),
          (4, 350), (8, 350), (8, 378), (8, 126), (8, 500), (4, 294), (4, 432), (8, 147), (8, 375), (8, 336), (4, 400), (8, 400)]
CASES = [("c2c", 8, 490), ("c2c", 8, 486), ("c2c", 8, 160), ("c2c", 8, 392), ("c2c", 4, 225), ("c2c", 4, 490),
         ("c2c", 4, 343), ("r2c", 8, 315), ("r2c", 4, 441), ("r2c", 4, 343), ("c2r", 4, 200), ("c2r", 8, 500)]


def wider(n, fp, ttype):
    full = 128 // (2 * fp)
    nn = n // 2 if (ttype != "c2c" and n % 2 == 0) else n
    out = []
    facs = []
    for L, max_r in ((2, 36 if fp == 4 else 27), (3, 18), (4, 8)):
        f = [x for x in tune_gpu.factorizations(nn, max_r, L) if len(x) == L]
        f.sort(key=lambda x: (max(x), sum(x)))
        facs += f[:4]
    perms = []
    for f in facs:
        perms.append(sorted(f))
        perms.append(sorted(f, reverse=True))
        if len(f) == 3:
            perms.append([f[1], f[0], f[2]])
    seen = set()
    for f in perms:
        ts = set()
        for r in f:
            ts.add(nn // r)
            ts.add(-(-(nn // r) // 2))
            ts.add(-(-(nn // r) // 3))
        for t in sorted(ts):
            if t < 1:
                continue
            elems = max(-(-(nn // r) // t) * r for r in f)
            if elems > (40 if fp == 4 else 28):
                continue
            for ml in {full, max(2, full // 2), max(2, full // 4)} if ttype == "c2c" else {full}:
                if ml * t > 1024:
                    continue
                for bh in (1, 2):
                    for mb in (1, 2, 3, 4):
                        thr = ml * t * bh
                        if thr > 1024 or thr * mb > 2048 or thr < 32:
                            continue
                        if ml * bh * nn * 2 * fp * mb > 220 * 1024:
                            continue
                        s = "R=%s,T=%d,ML=%d,BH=%d,MB=%d" % ("x".join(map(str, f)), t, ml, bh, mb)
                        if s not in seen:
                            seen.add(s)
                            out.append(s)
    return out


def main():
    cases = {}
    todo = CASES
    cap = 420
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        todo = [("c2c", fp, n) for fp, n in ROUND2]
        cap = 150
    for ttype, fp, n in todo:
        elem = fp if ttype == "r2c" else 2 * fp
        k = max(2, (1 << 30) // (16 * n * elem)) // 2 * 2
        if ttype == "c2r":
            k = max(2, (1 << 30) // (16 * n * fp)) // 2 * 2
        desc = "%s%s%so16.%d*%d" % ("s" if fp == 4 else "d", "c" if ttype == "c2c" else "r", "b" if ttype == "c2r" else "f", n, k)
        cands = [""] + [c for c in tune_gpu.candidates(n, fp, 16, ttype) if c]
        extra = [c for c in wider(n, fp, ttype) if c not in cands]
        import random
        random.Random(n * 8 + fp).shuffle(extra)
        cases[desc] = cands + extra[: max(0, cap - len(cands))]
    json.dump(cases, sys.stdout, indent=0)
    sys.stderr.write("%s\n" % {d: len(c) for d, c in cases.items()})


if __name__ == "__main__":
    main()
