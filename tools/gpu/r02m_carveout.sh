#!/bin/bash
# round 2, pass m: shared-memory carve-out preference on/off over the sweep, twice each, alternating
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02m_carveout.log
: > $L
for rep in 1 2; do
for v in 0 1; do
  BBFFT_CUDA_CARVEOUT=$v timeout 300 python bench.py --steps 4 --warmup 3 --no-extra --e2e-steps 0 --no-cpu-baseline \
      --per-size $OUT/r02m_per_size_c${v}_$rep.csv > $OUT/r02m_bench_c${v}_$rep.json 2>> $OUT/r02m.err
  python - <<PY >> $L
import json
d=json.load(open("$OUT/r02m_bench_c${v}_$rep.json"))
r=d["roofline"]
print("carveout=$v rep=$rep value=%.0f frac=%.4f min=%.3f n<0.8=%d n<0.85=%d below=%s" % (d["value"], r["frac"], r["per_size_frac"]["min"], r["per_size_frac"]["n_below_0.8"], r["per_size_frac"]["n_below_0.85"], r["below_0.8"]))
PY
done
done
cat $L | cut -c1-400
