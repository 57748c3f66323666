#!/bin/bash
# round 2, pass y: the fused real tile parity test again (imaginary parts of the real spectrum entries ignored), then
# packed fp32 adds per kernel over the fp32 c2c sweep and the fp32 r2c / c2r M=16 sweep
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_nd.py -x -q -m gpu -k "real_nd_fused or persistent_tile" 2>&1 | tail -4
BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0 timeout 900 python tools/bench_tile_ab.py --which x2sweep --rounds 6 > $OUT/r02y_x2sweep.log 2> $OUT/r02y.err
tail -3 $OUT/r02y.err; awk '{ if ($9+0 > 1.02 || $9+0 < 0.98) print }' $OUT/r02y_x2sweep.log | cut -c1-120; wc -l $OUT/r02y_x2sweep.log
