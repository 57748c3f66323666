#!/bin/bash
# callback/r2c parity after the explicit-FMA change, BASELINE configs C1/C3/C4/C5, then the re-tune
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_callback.py tests/test_gpu_r2c.py tests/test_gpu_c2c.py -x -q > $OUT/p2_pytest.log 2>&1; tail -3 $OUT/p2_pytest.log
timeout 900 python tools/bench_configs.py --out $OUT/p2_configs.csv > $OUT/p2_configs.jsonl 2> $OUT/p2_configs.err; tail -3 $OUT/p2_configs.err
BBFFT_CUDA_NO_WISDOM=1 timeout 2400 python tools/tune_gpu.py --minN 6 --out $OUT/wisdom2_c2c_m16.json > $OUT/p2_tune.log 2>&1
tail -3 $OUT/p2_tune.log
