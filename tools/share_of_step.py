"""Share of the step per kernel: the ncu launch list of the bench command (cold-cache, serialised launches)
against the per-launch CUDA-event times of the same bench (per-size CSV).  The absolute times differ by
design; every kernel's SHARE of the step must agree.
Usage: python tools/share_of_step.py <ncu_launches.csv> <per_size.csv> > profiles/<tag>_share_of_step.txt"""
import csv
import sys


def main():
    ncu_path, per_size_path = sys.argv[1:3]
    ncu = {}
    rows = [r for r in csv.reader(open(ncu_path)) if len(r) > 10 and r[0].isdigit()]
    for r in rows:
        name, metric, val = r[4], r[12], r[14]
        if metric == "gpu__time_duration.sum" and name.startswith("bbfft_"):
            ncu.setdefault(name, []).append(float(val.replace(",", "")) * 1e-3)  # ns -> us
    ev = {}
    for r in csv.DictReader(open(per_size_path)):
        ev[r["kernel"]] = float(r.get("mean_time_us") or r["time_us"])
    # the launch list holds warm-up + timed steps: average the launches of every kernel
    ncu_avg = {k: sum(v) / len(v) for k, v in ncu.items()}
    both = sorted(set(ncu_avg) & set(ev))
    tn, te = sum(ncu_avg[k] for k in both), sum(ev[k] for k in both)
    worst = max(abs(ncu_avg[k] / tn - ev[k] / te) for k in both)
    wk = max(both, key=lambda k: abs(ncu_avg[k] / tn - ev[k] / te))
    print("# Share of the step per kernel: ncu launch list of the bench command (%s," % ncu_path)
    print("# `ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache and serialised) against the")
    print("# per-launch CUDA-event times of the same bench (%s)." % per_size_path)
    print("kernels in both lists: %d of %d" % (len(both), len(ev)))
    print("sum of per-launch durations: ncu %.1f ms, bench events %.1f ms" % (tn * 1e-3, te * 1e-3))
    print("largest difference in any kernel's share of the step: %.5f (absolute), kernel %s" % (worst, wk))
    top = sorted(both, key=lambda k: -ev[k])[:5]
    for k in top:
        print("  %-100s ncu %.4f  events %.4f" % (k[:100], ncu_avg[k] / tn, ev[k] / te))


if __name__ == "__main__":
    main()
