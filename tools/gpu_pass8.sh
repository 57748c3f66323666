#!/bin/bash
# pass 8: tile-kernel overrides, headline bench with the interleaved-regime wisdom, real sweep, nd tests
set -u
TAG=r01g
OUT=gpurun_out; mkdir -p $OUT
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0
timeout 300 python tools/tune_tile.py --out $OUT/${TAG}_tile_tune.json > $OUT/${TAG}_tile_tune.log 2>&1; cat $OUT/${TAG}_tile_tune.log | cut -c1-110
timeout 600 python tools/bench_configs.py --which none --real-sweep > $OUT/${TAG}_real_sweep.jsonl 2>&1; tail -1 $OUT/${TAG}_real_sweep.jsonl | cut -c1-200
unset BBFFT_CUDA_KERNEL_CACHE BBFFT_CUDA_JIT_LINEINFO
timeout 900 python bench.py --per-size $OUT/${TAG}_per_size.csv > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
timeout 600 python -m pytest tests/test_gpu_nd.py -x -q > $OUT/${TAG}_pytest_nd.log 2>&1; tail -3 $OUT/${TAG}_pytest_nd.log
ls $OUT | grep $TAG
