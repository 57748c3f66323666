#!/bin/bash
# round 2, pass f: tile kernel split over a thread-block cluster (DSMEM) -- parity and A/B against one CTA per tile
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02f_cluster.log
: > $L
fmt='
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print("  %-22s %-16s %9.2f us %8.1f GB/s  cufft %s  err %.1e %s %s" % (r["config"], r["shape"], r["time_us"], r["GBs"], ("%.1f" % r["cufft_GBs"]) if r["cufft_GBs"] else "-", r["err"], r["note"], r["kernel"][-30:]))
'
echo "== nd tests" >> $L
timeout 900 python -m pytest tests/test_gpu_nd.py tests/test_gpu_examples.py -x -q -m gpu 2>&1 | tail -6 >> $L
for v in 0 1 2 4 8; do
  echo "== C4, BBFFT_CUDA_TILE_CLUSTER=$v (0 = planner's choice)" >> $L
  if [ $v = 0 ]; then unset BBFFT_CUDA_TILE_CLUSTER; else export BBFFT_CUDA_TILE_CLUSTER=$v; fi
  timeout 600 python tools/bench_configs.py --which c4 2>> $OUT/r02f.err | python -c "$fmt" >> $L
done
unset BBFFT_CUDA_TILE_CLUSTER
echo "== spill probe: real sweep laggards + C1 burst (PDL default)" >> $L
timeout 900 python tools/bench_configs.py --which none --real-sweep 2>> $OUT/r02f.err > $OUT/r02f_real_sweep.jsonl
python - <<PY >> $L
import json
rows=[json.loads(l) for l in open("$OUT/r02f_real_sweep.jsonl") if l.startswith("{")]
peak=6534.5
fr=sorted((r["GBs"]/peak, r["config"], r["fp"], r["shape"], r["kernel"]) for r in rows)
print("  rows=%d min=%.3f median=%.3f n<0.8=%d" % (len(fr), fr[0][0], fr[len(fr)//2][0], sum(1 for f in fr if f[0]<0.8)))
for f in fr[:10]: print("   ", round(f[0],3), f[1], f[2], f[3], f[4])
PY
timeout 120 tools/bin/bbfft-bench -o -m 1 -k 16384 --burst 200 --impl bbfft sc 64 >> $L 2>&1
cat $L | cut -c1-330
