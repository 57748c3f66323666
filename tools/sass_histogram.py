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
KEY = ["LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "SYNCS", "UBLKPF", "UTMALDG", "UTMASTG", "SHFL", "BAR", "UCGABAR_ARV", "UCGABAR_WAIT",
       "ACQBULK", "LDL", "STL", "FFMA", "FADD", "FADD2", "FMUL", "DFMA", "DADD", "DMUL", "IMAD", "HMMA", "UTCMMA", "LDC", "LDCU"]
# plans of BASELINE config 4 and its real twins whose default kernel is JIT-compiled (not in the bundle)
JIT_DEFAULTS = [("scfo128x128*8192", "c2c2d staged+bulk"), ("srfo128x128*8192", "r2c2d fused real"), ("srbo128x128*8192", "c2r2d fused real")]


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
    sys.path.insert(0, ROOT)
    import importlib
    import tempfile
    pkg = importlib.import_module("double-batched-fft-library_b200")
    for desc, label in JIT_DEFAULTS:
        d = pkg.describe(pkg.parse_descriptor(desc))
        with tempfile.NamedTemporaryFile(suffix=".cubin") as tf:
            tf.write(pkg.compile_to_cubin(d["source"]))
            tf.flush()
            out = subprocess.run(["cuobjdump", "-sass", tf.name], capture_output=True, text=True).stdout
        f = "jit: " + label
        nk[f] += 1
        for line in out.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                fam[f][m.group(1)] += 1
    print("# static SASS opcode counts of the built-in bundle (nvcc -gencode arch=compute_100a,code=sm_100a), %d kernels" % sum(nk.values()))
    print("# memory path: LDG/STG (global), LDS/STS (shared), BAR (CTA barrier), UBLKPF (cp.async.bulk.prefetch.L2: the")
    print("# prefetch switch, executed only when args.pf != 0), ACQBULK / griddepcontrol (PDL prologue).  The 1d kernels use")
    print("# no TMA, no LDGSTS, no SHFL and no tensor-core instructions (HMMA/UTCMMA): they saturate HBM with plain loads.")
    print("# FADD2 = packed fp32 adds (add.f32x2) in the kernels whose wisdom entry carries X2=1.  The rows 'jit:' are the")
    print("# DEFAULT plans of the 128 x 128 fp32 tiles, compiled at plan creation: the staged persistent c2c tile kernel fills")
    print("# its staging buffer with cp.async.bulk (UBLKCP) completing on an mbarrier (SYNCS.*) and the rest of the tile with")
    print("# cp.async (LDGSTS); the fused real tiles are one-tile-per-CTA kernels.  The PS=1 pipeline and the cluster/DSMEM")
    print("# tile kernel (CL>1) stay switches that measured slower (profiles/r02e_tile_pdl.txt, r02f_cluster.txt).")
    print("%-24s %8s  %s" % ("family", "kernels", "  ".join("%s" % k for k in KEY)))
    for f in sorted(fam):
        print("%-24s %8d  %s" % (f, nk[f], "  ".join("%*d" % (len(k), fam[f][k]) for k in KEY)))
    print("%-24s %8d  %s" % ("TOTAL (bundle)", sum(v for k, v in nk.items() if not k.startswith("jit")), "  ".join("%*d" % (len(k), total[k]) for k in KEY)))
    print("\n# all opcodes, total:")
    print(", ".join("%s %d" % kv for kv in total.most_common()))


if __name__ == "__main__":
    main()
