#!/bin/bash
# pass 9: tuner again with HBM-saturating filler launches between candidates (the sweep's power /
# clock regime), more tile-kernel overrides, chained 3d with the 4x16 tile
set -u
TAG=r01h
OUT=gpurun_out; mkdir -p $OUT
[ -f tune_cache_gpu.tar.xz ] && tar -xJf tune_cache_gpu.tar.xz
CANDS=$(ls tune_cache_gpu/cands_*.json 2>/dev/null | paste -sd, -)
NCUBIN=$(ls tune_cache_gpu 2>/dev/null | grep -c cubin)
echo "pre-compiled candidates on the box: $NCUBIN"
if [ -n "$CANDS" ] && [ "$NCUBIN" -gt 12000 ]; then
  BBFFT_CUDA_KERNEL_CACHE=tune_cache_gpu BBFFT_CUDA_JIT_LINEINFO=0 BBFFT_CUDA_NO_WISDOM=1 timeout 1500 \
    python tools/tune_gpu.py --cands "$CANDS" --filler 2 --out $OUT/${TAG}_wisdom.json > $OUT/${TAG}_tune.log 2>&1
  tail -2 $OUT/${TAG}_tune.log | cut -c1-200
fi
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0
timeout 300 python tools/tune_tile.py --out $OUT/${TAG}_tile_tune.json > $OUT/${TAG}_tile_tune.log 2>&1; cat $OUT/${TAG}_tile_tune.log | cut -c1-110
unset BBFFT_CUDA_KERNEL_CACHE BBFFT_CUDA_JIT_LINEINFO
timeout 600 python tools/bench_configs.py --which c4 > $OUT/${TAG}_c4.jsonl 2>&1; cut -c1-160 $OUT/${TAG}_c4.jsonl | tail -12
ls $OUT | grep $TAG
