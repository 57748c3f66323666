"""Convert the auto-tuner's JSON (tools/tune_gpu.py) into csrc/wisdom.inc, the table of measured
planner overrides that is compiled into the library.  Entries that do not beat the heuristic
default by more than `--min-gain` are dropped.
Usage: python tools/make_wisdom.py gpurun_out/wisdom_c2c_m16.json [more.json ...]"""
import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="+")
    ap.add_argument("--min-gain", type=float, default=0.03)
    ap.add_argument("--no-pin", action="store_true", help="keep uncapped (MB=1) entries as measured")
    ap.add_argument("--out", default=os.path.join(ROOT, "double-batched-fft-library_b200", "csrc", "wisdom.inc"))
    ap.add_argument("--real-out", default=os.path.join(ROOT, "double-batched-fft-library_b200", "csrc", "wisdom_real.inc"))
    ap.add_argument("--keep-existing", action="store_true", help="merge into the current wisdom.inc instead of rebuilding it")
    ap.add_argument("--min-switch", type=float, default=0.015, help="gain needed to replace a current entry that was re-measured")
    ap.add_argument("--real-only", action="store_true", help="only (re)write wisdom_real.inc; leave wisdom.inc alone")
    args = ap.parse_args()
    entries = {}
    real_entries = {}
    if args.keep_existing and os.path.exists(args.out):
        # start from the table that is compiled in today; the given files supersede it size by size
        import re as _re
        for line in open(args.out):
            m = _re.match(r'\{(\d+), (\d+), "([^"]*)"\}, // (\d+) GB/s \(heuristic (\S+)\)', line)
            if m:
                dflt = None if m.group(5) == "n/a" else float(m.group(5))
                entries[(int(m.group(1)), int(m.group(2)))] = (m.group(3), float(m.group(4)), dflt)
    if args.keep_existing and os.path.exists(args.real_out):
        import re as _re
        for line in open(args.real_out):
            m = _re.match(r'\{(\d+), (\d+), (\d+), 0, "([^"]*)"\}, // (\d+) GB/s \(heuristic (\S+)\)', line)
            if m:
                dflt = None if m.group(6) == "n/a" else float(m.group(6))
                real_entries[(int(m.group(1)), int(m.group(2)), int(m.group(3)))] = (m.group(4), float(m.group(5)), dflt)
    for f in args.files:
        for key, v in json.load(open(f)).items():
            parts = key.split(",")
            if len(parts) == 3:
                # "r2c,fp,N" / "c2r,fp,N": measured real-transform entries go to wisdom_real.inc
                tkey = ({"r2c": 1, "c2r": 2}[parts[0]], int(parts[1]), int(parts[2]))
                if args.keep_existing and tkey in real_entries and v.get("mode") == "interleaved":
                    cur = real_entries[tkey][0]
                    cur_gbs = dict((t, g) for t, g in v.get("top", [])).get(cur)
                    if v["best"] == cur or (cur_gbs and v["gbs"] < cur_gbs * (1.0 + args.min_switch)):
                        continue
                real_entries.pop(tkey, None)
                if v["best"] and not (v.get("default_gbs") and v["gbs"] < v["default_gbs"] * (1.0 + args.min_gain)):
                    real_entries[tkey] = (v["best"], v["gbs"], v.get("default_gbs"))
                continue
            fp, n = [int(x) for x in parts]
            if args.keep_existing and (fp, n) in entries and v.get("mode") == "interleaved":
                # the current entry competed as a candidate: keep it unless the winner is clearly better
                cur = entries[(fp, n)][0]
                cur_gbs = dict((t, g) for t, g in v.get("top", [])).get(cur)
                if v["best"] == cur or (cur_gbs and v["gbs"] < cur_gbs * (1.0 + args.min_switch)):
                    continue
            # a later file supersedes whatever an earlier one said about this size
            entries.pop((fp, n), None)
            if not v["best"]:
                continue
            if v.get("default_gbs") and v["gbs"] < v["default_gbs"] * (1.0 + args.min_gain):
                continue
            entries[(fp, n)] = (v["best"], v["gbs"], v.get("default_gbs"))
    # An entry without a register cap (MB=1 or no MB) was measured at whatever occupancy NVRTC's
    # register count happened to give; pin that occupancy so that the nvcc-built bundle (which
    # lands a few registers higher or lower on the same source) runs at the measured one.
    if not args.no_pin:
        import re
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import occupancy_table as ot
        for (fp, n), (best, gbs, dflt) in list(entries.items()):
            m = re.search(r"MB=(\d+)", best)
            if m and int(m.group(1)) > 1:
                continue
            desc = "%scfo16.%d*%d" % ("s" if fp == 4 else "d", n, ot.aot.sweep_k(n, fp))
            _, v = ot.one(desc, best)
            if v is None or v["occ_cons"] < 2:
                continue
            pinned = re.sub(r",?MB=\d+", "", best) + ",MB=%d" % min(v["occ_cons"], 8)
            _, w = ot.one(desc, pinned)
            if w is not None and w["stack"] <= v["stack"]:
                entries[(fp, n)] = (pinned, gbs, dflt)
    if real_entries or args.real_only:
        rl = ["// generated by tools/make_wisdom.py from B200 measurements (tools/tune_gpu.py --type r2c|c2r); do not edit",
              "// {transform type (1 = r2c, 2 = c2r), bytes per real, N, 0 (= any M filling whole rows), planner overrides}"]
        for (ty, fp, n) in sorted(real_entries):
            best, gbs, dflt = real_entries[(ty, fp, n)]
            rl.append('{%d, %d, %d, 0, "%s"}, // %.0f GB/s (heuristic %s)' % (ty, fp, n, best, gbs, "%.0f" % dflt if dflt else "n/a"))
        with open(args.real_out, "w") as f:
            f.write("\n".join(rl) + "\n")
        print("wrote %d entries to %s" % (len(real_entries), args.real_out))
        if args.real_only:
            return
    lines = ["// generated by tools/make_wisdom.py from B200 measurements (tools/tune_gpu.py); do not edit",
             "// {bytes per real, N, planner overrides}   -- 1d c2c, batch lanes full (M multiple of ML)"]
    for (fp, n) in sorted(entries):
        best, gbs, dflt = entries[(fp, n)]
        lines.append('{%d, %d, "%s"}, // %.0f GB/s (heuristic %s)' % (fp, n, best, gbs, "%.0f" % dflt if dflt else "n/a"))
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote %d entries to %s" % (len(entries), args.out))


if __name__ == "__main__":
    main()
