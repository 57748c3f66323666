#!/bin/bash
# round 2, pass t: staged persistent tile kernel (bbk::fft2d_tile_staged) -- parity, then A/B on config 4
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02t_staged.log
: > $L
echo "== parity (persistent tile pipelines)" >> $L
timeout 900 python -m pytest tests/test_gpu_nd.py -x -q -m gpu -k "persistent_tile_pipelines" 2>&1 | tail -4 >> $L
echo "== A/B" >> $L
BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0 timeout 900 python tools/bench_tile_ab.py >> $L 2>> $OUT/r02t.err
cat $L | cut -c1-300
