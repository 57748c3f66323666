#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_nd.py tests/test_gpu_r2c.py tests/test_gpu_callback.py -x -q > $OUT/p4_pytest.log 2>&1; tail -3 $OUT/p4_pytest.log
timeout 900 python tools/bench_configs.py --which c1,c3 --real-sweep > $OUT/p4_real_sweep.jsonl 2>&1
BBFFT_CUDA_NO_WISDOM=1 timeout 2400 python tools/tune_gpu.py --from-csv profiles/r01c_per_size.csv --below 0.9 --out $OUT/wisdom3_sustained.json > $OUT/p4_tune.log 2>&1
tail -3 $OUT/p4_tune.log
