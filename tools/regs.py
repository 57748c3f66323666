"""Compile one planned kernel with NVRTC (no GPU needed) and print registers / spills / smem.
Usage: python tools/regs.py <descriptor> [tune] ..."""
import importlib, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("double-batched-fft-library_b200")

def usage(desc, tune=""):
    cfg = pkg.parse_descriptor(desc)
    d = pkg.describe(cfg, tune)
    cubin = pkg.compile_to_cubin(d["source"])
    with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
        f.write(cubin); f.flush()
        out = subprocess.run(["cuobjdump", "-res-usage", f.name], capture_output=True, text=True).stdout
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out)
    reg, stack, shared, local = map(int, m.groups())
    thr = d["threads"]
    ra = (reg + 7) // 8 * 8
    by_reg = 65536 // (ra * ((thr + 31) // 32 * 32))
    by_smem = (228 * 1024) // (d["smem_bytes"] + 1024) if d["smem_bytes"] else 32
    occ = min(by_reg, by_smem, 2048 // thr, 32)
    return dict(id=d["identifier"], reg=reg, stack=stack, local=local, threads=thr, smem=d["smem_bytes"], ctas=occ, warps=occ * ((thr + 31) // 32))

if __name__ == "__main__":
    desc = sys.argv[1]
    for tune in (sys.argv[2:] or [""]):
        u = usage(desc, tune)
        print("%-40s reg %3d stack %3d thr %4d smem %6d CTAs/SM %d warps %d" % (tune, u["reg"], u["stack"], u["threads"], u["smem"], u["ctas"], u["warps"]))
