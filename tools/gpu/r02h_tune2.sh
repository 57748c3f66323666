#!/bin/bash
# round 2, pass h: second override search (sweep sizes between 0.80 and 0.90) + the sweep with the first search's winners
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02h_tune2.log
: > $L
timeout 300 python bench.py --steps 4 --warmup 3 --no-extra --e2e-steps 0 --no-cpu-baseline --per-size $OUT/r02h_per_size.csv > $OUT/r02h_bench.json 2>> $OUT/r02h.err
python - <<PY >> $L
import json
d=json.load(open("$OUT/r02h_bench.json"))
r=d["roofline"]
print("value=%.0f GFLOP/s frac=%.4f min=%.3f n<0.8=%d n<0.85=%d below=%s" % (d["value"], r["frac"], r["per_size_frac"]["min"], r["per_size_frac"]["n_below_0.8"], r["per_size_frac"]["n_below_0.85"], r["below_0.8"]))
PY
timeout 1500 python tools/tune_list.py --cases tools/cases_laggards2.json --reps 5 --out $OUT/r02h_laggards2.json > $OUT/r02h_laggards2.log 2>&1
grep -c rejected $OUT/r02h_laggards2.log >> $L
grep -v rejected $OUT/r02h_laggards2.log | awk '{c[$1]++; if (c[$1]<=3) print}' >> $L
cat $L | cut -c1-230
