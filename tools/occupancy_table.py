"""For every kernel of the C2 sweep: compile the planned stub with NVRTC (as the tuner did) and
record registers and the occupancy they imply under the per-SM-sub-partition register model.
Usage: python tools/occupancy_table.py out.json"""
import importlib, json, os, re, subprocess, sys, tempfile
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("double-batched-fft-library_b200")
aot = importlib.import_module("double-batched-fft-library_b200.aot")


def real_occupancy(reg, threads, smem):
    warps = (threads + 31) // 32
    per_warp = ((reg + 7) // 8 * 8) * 32
    warps_per_smsp = 16384 // per_warp          # register file is split over 4 sub-partitions
    by_reg = (warps_per_smsp * 4) // warps
    # each CTA's warps are dealt round-robin over the sub-partitions starting at 0 (conservative)
    by_reg_cons = warps_per_smsp // ((warps + 3) // 4)
    by_smem = (228 * 1024) // (smem + 1024) if smem else 32
    return min(by_reg, by_smem, 2048 // threads, 32), min(by_reg_cons, by_smem, 2048 // threads, 32)


def one(desc, tune=""):
    cfg = pkg.parse_descriptor(desc)
    try:
        d = pkg.describe(cfg, tune)
    except pkg.BadConfiguration:
        return desc, None  # multi-kernel (nd) plan
    cubin = pkg.compile_to_cubin(d["source"])
    with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
        f.write(cubin); f.flush()
        out = subprocess.run(["cuobjdump", "-res-usage", f.name], capture_output=True, text=True).stdout
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out)
    reg, stack, shared, local = map(int, m.groups())
    occ, occ_cons = real_occupancy(reg, d["threads"], d["smem_bytes"])
    return desc, dict(id=d["identifier"], reg=reg, stack=stack, local=local, threads=d["threads"],
                      smem=d["smem_bytes"], occ=occ, occ_cons=occ_cons)


if __name__ == "__main__":
    descs = [d for d in aot.BUILTIN_DESCRIPTORS]
    with ThreadPoolExecutor(8) as ex:
        res = {k: v for k, v in ex.map(one, descs) if v is not None}
    json.dump(res, open(sys.argv[1], "w"), indent=1)
    for k, v in res.items():
        mb = int(re.search(r"_mb(\d+)_", v["id"]).group(1))
        flag = "" if v["occ"] == v["occ_cons"] else "  <-- models differ"
        print("%-22s reg %3d thr %4d smem %6d mb %d occ %d/%d%s" % (k, v["reg"], v["threads"], v["smem"], mb, v["occ"], v["occ_cons"], flag))
