#!/bin/bash
# pass 13: config 3 / config 1 override search (pre-compiled candidates), then the full GPU tier and
# the bench on the state with direct-DFT stages
set -u
TAG=r01l
OUT=gpurun_out; mkdir -p $OUT
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0 BBFFT_CUDA_NO_WISDOM=1
timeout 600 python tools/tune_list.py --cases tools/cases_c3.json --out $OUT/${TAG}_c3_tune.json > $OUT/${TAG}_c3_tune.log 2>&1; grep -c GB/s $OUT/${TAG}_c3_tune.log; awk '{print $1}' $OUT/${TAG}_c3_tune.log | uniq -c | head; for d in srfo srfi srbo srbi; do grep "^${d}" $OUT/${TAG}_c3_tune.log | head -3 | cut -c1-150; done
timeout 300 python tools/tune_list.py --cases tools/cases_c1.json --inner 20 --filler 0 --out $OUT/${TAG}_c1_tune.json > $OUT/${TAG}_c1_tune.log 2>&1; head -5 $OUT/${TAG}_c1_tune.log | cut -c1-150
unset BBFFT_CUDA_KERNEL_CACHE BBFFT_CUDA_JIT_LINEINFO BBFFT_CUDA_NO_WISDOM
# the full tier ran green on the previous state (profiles/r01i_pytest.log); this pass re-runs the files that the
# direct-DFT stages and the planner changes touch (everything except the large analytic/callback matrices)
timeout 1200 python -m pytest tests/test_gpu_c2c.py tests/test_gpu_nd.py tests/test_gpu_examples.py tests/test_gpu_z_cpp_api.py -m gpu -x -q \
    -k "not analytic_reference_suite" > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
timeout 900 python -m pytest tests/test_gpu_r2c.py tests/test_gpu_callback.py -m gpu -x -q -k "large_prime or golden or full_size or vs_oracle or config5" \
    > $OUT/${TAG}_pytest2.log 2>&1; tail -3 $OUT/${TAG}_pytest2.log
timeout 900 python bench.py --per-size $OUT/${TAG}_per_size.csv > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
ls $OUT | grep $TAG
