#!/bin/bash
# round 2, pass e: persistent async tile kernel A/B, PDL on config 1, real sweep with/without L2 prefetch,
# one full bench line (e2e ring + copy ceiling)
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02e_tile_pdl.log
: > $L
fmt='
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print("  %-22s %-16s %9.2f us %8.1f GB/s  cufft %s  err %.1e %s" % (r["config"], r["shape"], r["time_us"], r["GBs"], ("%.1f" % r["cufft_GBs"]) if r["cufft_GBs"] else "-", r["err"], r["note"]))
'
echo "== nd tests" >> $L
timeout 600 python -m pytest tests/test_gpu_nd.py tests/test_gpu_examples.py -x -q -m gpu 2>&1 | tail -4 >> $L
for v in 1 0; do
  echo "== C4, BBFFT_CUDA_TILE_ASYNC=$v" >> $L
  BBFFT_CUDA_TILE_ASYNC=$v timeout 600 python tools/bench_configs.py --which c4 2>> $OUT/r02e.err | python -c "$fmt" >> $L
done
for v in 0 1; do
  echo "== C1, BBFFT_CUDA_PDL=$v" >> $L
  BBFFT_CUDA_PDL=$v timeout 600 python tools/bench_configs.py --which c1 2>> $OUT/r02e.err | python -c "$fmt" >> $L
  BBFFT_CUDA_PDL=$v timeout 120 tools/bin/bbfft-bench -o -m 1 -k 16384 --burst 200 --impl bbfft sc 64 >> $L 2>&1
  BBFFT_CUDA_PDL=$v timeout 120 tools/bin/bbfft-bench -o -m 16 --burst 20 --impl bbfft sc 64 256 >> $L 2>&1
done
for w in 0 1.5; do
  echo "== real sweep M=16, BBFFT_CUDA_PREFETCH_WAVES=$w" >> $L
  BBFFT_CUDA_PREFETCH_WAVES=$w timeout 900 python tools/bench_configs.py --which none --real-sweep 2>> $OUT/r02e.err > $OUT/r02e_real_sweep_w$w.jsonl
  python - <<PY >> $L
import json
rows=[json.loads(l) for l in open("$OUT/r02e_real_sweep_w$w.jsonl") if l.startswith("{")]
peak=6534.5
fr=sorted((r["GBs"]/peak, r["config"], r["fp"], r["shape"]) for r in rows)
print("  rows=%d min=%.3f median=%.3f n<0.8=%d  lowest: %s" % (len(fr), fr[0][0], fr[len(fr)//2][0], sum(1 for f in fr if f[0]<0.8), [(round(f[0],3),f[1],f[2],f[3]) for f in fr[:8]]))
PY
done
echo "== bench.py full line" >> $L
timeout 1200 python bench.py --steps 5 --warmup 3 --per-size $OUT/r02e_per_size.csv > $OUT/r02e_bench.json 2>> $OUT/r02e.err
python - <<PY >> $L
import json
d=json.load(open("$OUT/r02e_bench.json"))
print("value %.0f frac %.4f e2e %s" % (d["value"], d["roofline"]["frac"], json.dumps(d["e2e"])[:600]))
print("cpu_baseline", json.dumps(d["cpu_baseline"])[:500])
print("sharded", json.dumps(d["sharded_configs"])[:3000])
print("below", d["roofline"]["below_0.8"])
PY
cat $L | cut -c1-700
