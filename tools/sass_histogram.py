"""Opcode histogram of the built-in kernel bundle (cuobjdump -sass of builtin_kernels_*.cubin): which memory /
synchronisation / arithmetic instructions the shipped kernels are made of, per kernel family and in total.
Usage: python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKPF", "UTMALDG", "UTMASTG", "SHFL", "BAR", "UCGABAR_ARV", "UCGABAR_WAIT", "ACQBULK",
       "LDL", "STL", "FFMA", "FADD", "FMUL", "DFMA", "DADD", "DMUL", "IMAD", "HMMA", "UTCMMA", "LDC", "LDCU"]


def main():
    total = collections.Counter()
    fam = collections.defaultdict(collections.Counter)
    nk = collections.Counter()
    for cubin in sorted(glob.glob(os.path.join(ROOT, "double-batched-fft-library_b200", "builtin_kernels_*.cubin"))):
        out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
        name = None
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                name = m.group(1)
                f = re.match(r"bbfft_([a-z0-9]+)_", name).group(1) + ("_f32" if "_f32_" in name else "_f64")
                nk[f] += 1
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m and name:
                op = m.group(1)
                total[op] += 1
                fam[f][op] += 1
    print("# static SASS opcode counts of the built-in bundle (nvcc -gencode arch=compute_100a,code=sm_100a), %d kernels" % sum(nk.values()))
    print("# memory path: LDG/STG (global), LDS/STS (shared), BAR (CTA barrier), UBLKPF (cp.async.bulk.prefetch.L2: the")
    print("# prefetch switch, executed only when args.pf != 0), ACQBULK / griddepcontrol (PDL prologue).  No TMA tensor")
    print("# copies (UTMALDG/UTMASTG), no LDGSTS, no SHFL, no tensor-core instructions (HMMA/UTCMMA) in the shipped kernels:")
    print("# the persistent cp.async tile kernel (PS=1) and the cluster/DSMEM tile kernel (CL>1) are switches that measured")
    print("# slower (profiles/r02e_tile_pdl.txt, r02f_cluster.txt) and are JIT-compiled only on request.")
    print("%-14s %8s  %s" % ("family", "kernels", "  ".join("%s" % k for k in KEY)))
    for f in sorted(fam):
        print("%-14s %8d  %s" % (f, nk[f], "  ".join("%*d" % (len(k), fam[f][k]) for k in KEY)))
    print("%-14s %8d  %s" % ("TOTAL", sum(nk.values()), "  ".join("%*d" % (len(k), total[k]) for k in KEY)))
    print("\n# all opcodes, total:")
    print(", ".join("%s %d" % kv for kv in total.most_common()))


if __name__ == "__main__":
    main()
