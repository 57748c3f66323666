#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
BBFFT_CUDA_NO_WISDOM=1 timeout 1500 python tools/tune_gpu.py --from-csv tools/retune_list.csv --below 2.0 --out $OUT/wisdom4_sustained.json > $OUT/p5_tune.log 2>&1
tail -2 $OUT/p5_tune.log
timeout 300 python -m pytest tests/test_gpu_nd.py -x -q 2>&1 | tail -2
