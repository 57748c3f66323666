#!/bin/bash
# round 2, pass u: packed fp32 adds (add.f32x2 -> FADD2) A/B on the tile kernel and on fp32 1d kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0 timeout 900 python tools/bench_tile_ab.py --which x2 > $OUT/r02u_x2.log 2> $OUT/r02u.err
cut -c1-250 $OUT/r02u_x2.log; tail -5 $OUT/r02u.err
