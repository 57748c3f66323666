#!/bin/bash
# round 2, pass o: the sweep with the mixed nvcc / NVRTC bundle and the in-sweep winners, three times
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02o_check.log
: > $L
for rep in 1 2 3; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --e2e-steps 0 --no-cpu-baseline \
      --per-size $OUT/r02o_per_size_$rep.csv > $OUT/r02o_bench_$rep.json 2>> $OUT/r02o.err
  python - <<PY >> $L
import json
d=json.load(open("$OUT/r02o_bench_$rep.json"))
r=d["roofline"]
print("rep=$rep value=%.0f frac=%.4f min=%.3f n<0.8=%d n<0.85=%d below=%s clocks=%s" % (d["value"], r["frac"], r["per_size_frac"]["min"], r["per_size_frac"]["n_below_0.8"], r["per_size_frac"]["n_below_0.85"], r["below_0.8"], d["clocks"]))
PY
done
cat $L | cut -c1-400
