#!/bin/bash
# pass 7: interleaved-regime tuning from pre-compiled candidates (c2c weak sizes + r2c/c2r sweep),
# first run of the chained nd kernel, real sweep with the new planner defaults
set -u
TAG=r01f
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_smi.txt 2>&1
[ -f tune_cache_gpu.tar.xz ] && tar -xJf tune_cache_gpu.tar.xz
CANDS=$(ls tune_cache_gpu/cands_*.json 2>/dev/null | paste -sd, -)
NCUBIN=$(ls tune_cache_gpu 2>/dev/null | grep -c cubin)
echo "pre-compiled candidates on the box: $NCUBIN" | tee $OUT/${TAG}_tune_cache.txt
if [ -n "$CANDS" ] && [ "$NCUBIN" -gt 12000 ]; then
  BBFFT_CUDA_KERNEL_CACHE=tune_cache_gpu BBFFT_CUDA_JIT_LINEINFO=0 BBFFT_CUDA_NO_WISDOM=1 timeout 1200 \
    python tools/tune_gpu.py --cands "$CANDS" --out $OUT/${TAG}_wisdom.json > $OUT/${TAG}_tune.log 2>&1
  tail -3 $OUT/${TAG}_tune.log
fi
timeout 600 python -m pytest tests/test_gpu_nd.py -x -q > $OUT/${TAG}_pytest_nd.log 2>&1; tail -3 $OUT/${TAG}_pytest_nd.log
timeout 600 python -m pytest tests/test_gpu_r2c.py -x -q -k "golden or full_size" > $OUT/${TAG}_pytest_r2c.log 2>&1; tail -3 $OUT/${TAG}_pytest_r2c.log
timeout 600 python tools/bench_configs.py --which c3,c4 > $OUT/${TAG}_c3c4.jsonl 2>&1; cut -c1-260 $OUT/${TAG}_c3c4.jsonl | tail -22

timeout 300 ncu --set full --clock-control none --import-source on -k regex:bbfft_chain -c 1 --launch-skip 2 -f -o $OUT/${TAG}_full_chain3d \
    python tools/bench_configs.py --which c4 > $OUT/${TAG}_full_chain3d.log 2>&1
ls $OUT | grep $TAG
