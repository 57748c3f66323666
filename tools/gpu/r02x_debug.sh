#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0
timeout 300 python tools/debug_staged.py > $OUT/r02x_debug.log 2>&1
echo "== PDL=0" >> $OUT/r02x_debug.log
BBFFT_CUDA_PDL=0 timeout 300 python tools/debug_staged.py >> $OUT/r02x_debug.log 2>&1
cut -c1-250 $OUT/r02x_debug.log | tail -60
