#!/bin/bash
# round 2, pass n: the sweep with the nvcc-built bundle against the same kernels built by NVRTC (twice each)
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02n_jit_vs_aot.log
: > $L
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0
for rep in 1 2; do
for v in 0 1; do
  BBFFT_CUDA_NO_BUILTIN=$v timeout 600 python bench.py --steps 4 --warmup 3 --no-extra --e2e-steps 0 --no-cpu-baseline \
      --per-size $OUT/r02n_per_size_nobuiltin${v}_$rep.csv > $OUT/r02n_bench_${v}_$rep.json 2>> $OUT/r02n.err
  python - <<PY >> $L
import json
d=json.load(open("$OUT/r02n_bench_${v}_$rep.json"))
r=d["roofline"]
print("no_builtin=$v rep=$rep value=%.0f frac=%.4f min=%.3f n<0.8=%d n<0.85=%d below=%s" % (d["value"], r["frac"], r["per_size_frac"]["min"], r["per_size_frac"]["n_below_0.8"], r["per_size_frac"]["n_below_0.85"], r["below_0.8"]))
PY
done
done
cat $L | cut -c1-400
