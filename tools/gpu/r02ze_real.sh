#!/bin/bash
# round 2, pass ze: fused real tiles, split step without the group leaders -- parity, A/B table, ncu
set -u
OUT=gpurun_out; mkdir -p $OUT
T=r02ze
timeout 600 python -m pytest tests/test_gpu_nd.py -x -q -m gpu -k "real_nd_fused or r2c_c2r_nd" > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log
tail -3 $OUT/${T}_pytest.log
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0
timeout 600 python tools/bench_tile_ab.py --which real --rounds 5 > $OUT/${T}_real_tiles.log 2> $OUT/${T}.err
grep -v "TH=\|MB=" $OUT/${T}_real_tiles.log | cut -c1-130
export BBFFT_CUDA_JIT_LINEINFO=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bbfft_r2c2d -c 1 --launch-skip 2 -f -o $OUT/${T}_full_r2c_tile \
    python tools/bench_tile_ab.py --which prof2 > $OUT/${T}_full_tiles.log 2>&1
