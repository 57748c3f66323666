#!/bin/bash
# round 2, pass d: L2 prefetch of future batches -- A/B over the prefetch distance in the bench regime,
# the other BASELINE configs, and the new host-path / real-shape tests
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02d_prefetch.log
: > $L
for w in 0 1 1.5 2.5 4; do
  BBFFT_CUDA_PREFETCH_WAVES=$w timeout 300 python bench.py --steps 4 --warmup 3 --no-extra --e2e-steps 0 --no-cpu-baseline \
      --per-size $OUT/r02d_per_size_w$w.csv > $OUT/r02d_bench_w$w.json 2>> $OUT/r02d_bench.err
  python - <<PY >> $L
import json
d=json.load(open("$OUT/r02d_bench_w$w.json"))
r=d["roofline"]
print("waves=$w value=%.0f GFLOP/s frac=%.4f min=%.3f n<0.8=%d n<0.85=%d worst=%s below=%s" % (d["value"], r["frac"], r["per_size_frac"]["min"], r["per_size_frac"]["n_below_0.8"], r["per_size_frac"]["n_below_0.85"], r["worst"]["N"], r["below_0.8"]))
PY
done
for w in 0 1.5 3; do
  echo "== other configs, waves=$w" >> $L
  BBFFT_CUDA_PREFETCH_WAVES=$w timeout 600 python tools/bench_configs.py --which c1,c3,c4 2>> $OUT/r02d_bench.err | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print('  %-22s %-14s %9.2f us %8.1f GB/s  cufft %s  err %.1e %s' % (r['config'], r['shape'], r['time_us'], r['GBs'], ('%.1f' % r['cufft_GBs']) if r['cufft_GBs'] else '-', r['err'], r['note']))
" >> $L
done
echo "== pytest host path + new real shapes" >> $L
timeout 1200 python -m pytest tests/test_gpu_host_path.py "tests/test_gpu_r2c.py::test_real_small_radix_times_large_prime_vs_oracle" tests/test_gpu_nd.py -x -q -m gpu 2>&1 | tail -8 >> $L
cat $L | cut -c1-400
