#!/bin/bash
# round 2, pass w: fused real tile kernels (r2c / c2r 2d and 3d) -- parity, then A/B against one launch per mode
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02w_real.log
: > $L
echo "== parity" >> $L
timeout 1200 python -m pytest tests/test_gpu_nd.py -x -q -m gpu -k "real_nd_fused or r2c_c2r_nd or chain_matches or persistent_tile" 2>&1 | tail -15 >> $L
echo "== A/B" >> $L
BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0 timeout 900 python tools/bench_tile_ab.py --which real >> $L 2>> $OUT/r02w.err
cat $L | cut -c1-260; tail -5 $OUT/r02w.err
