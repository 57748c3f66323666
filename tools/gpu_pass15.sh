#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m pytest tests/test_gpu_z_cpp_api.py -m gpu -x -q -s > $OUT/r01n_cpp_api.log 2>&1; grep -E "callback|CHECK|OK|FAILED|passed|failed" $OUT/r01n_cpp_api.log | head -20 | cut -c1-250
