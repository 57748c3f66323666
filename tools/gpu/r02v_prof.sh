#!/bin/bash
# round 2, pass v: ncu --set full of the 128 x 128 fp32 tile kernel (plain / packed adds / staged bulk) and the
# packed-add A/B again with the strength-reduced tile indexing (everything JIT-built from the current header)
set -u
OUT=gpurun_out; mkdir -p $OUT
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=1 BBFFT_CUDA_NO_BUILTIN=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bbfft_c2c2d -c 24 -f -o $OUT/r02v_tile128 \
    python tools/bench_tile_ab.py --which prof > $OUT/r02v_prof.log 2>&1
tail -5 $OUT/r02v_prof.log
export BBFFT_CUDA_JIT_LINEINFO=0
timeout 900 python tools/bench_tile_ab.py --which x2 > $OUT/r02v_x2.log 2> $OUT/r02v.err
cut -c1-200 $OUT/r02v_x2.log; tail -5 $OUT/r02v.err
timeout 900 python tools/bench_tile_ab.py --which tile > $OUT/r02v_tile.log 2>> $OUT/r02v.err
cut -c1-200 $OUT/r02v_tile.log
