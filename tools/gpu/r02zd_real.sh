#!/bin/bash
# round 2, pass zd: fused real tiles with the power-of-two split mapping -- parity, A/B table, bench line, ncu
set -u
OUT=gpurun_out; mkdir -p $OUT
T=r02zd
timeout 900 python -m pytest tests/test_gpu_nd.py tests/test_gpu_examples.py -x -q -m gpu -k "real_nd_fused or r2c_c2r_nd or examples or persistent_tile" > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log
tail -3 $OUT/${T}_pytest.log
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0
timeout 600 python tools/bench_tile_ab.py --which real --rounds 5 > $OUT/${T}_real_tiles.log 2> $OUT/${T}.err
grep -v "TH=\|MB=" $OUT/${T}_real_tiles.log | cut -c1-130
unset BBFFT_CUDA_KERNEL_CACHE BBFFT_CUDA_JIT_LINEINFO
timeout 900 python bench.py --per-size $OUT/${T}_per_size.csv > $OUT/${T}_bench.json 2>> $OUT/${T}.err; echo "bench rc=$?"
cut -c1-200 $OUT/${T}_bench.json
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bbfft_r2c2d -c 1 --launch-skip 2 -f -o $OUT/${T}_full_r2c_tile \
    python tools/bench_tile_ab.py --which prof2 > $OUT/${T}_full_tiles.log 2>&1
ls -la $OUT | grep $T | cut -c1-100
