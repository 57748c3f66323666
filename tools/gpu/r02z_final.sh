#!/bin/bash
# round 2, pass z: final state -- full GPU tier, the bench line (+ per-size table), the reference arm, ncu launch list of
# the bench command, ncu --set full of the tile kernels (c2c staged, r2c fused), tile / real-tile A/B tables
set -u
OUT=gpurun_out; mkdir -p $OUT
T=${1:-r02z}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/${T}_smi.txt 2>&1
( time timeout 1700 python -m pytest tests -m gpu -x -q ) > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log
tail -6 $OUT/${T}_pytest.log
timeout 1500 python bench.py --per-size $OUT/${T}_per_size.csv > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"
cut -c1-400 $OUT/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > $OUT/${T}_bench_reference.json 2>> $OUT/${T}_bench.err; echo "reference rc=$?"
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0
timeout 600 python tools/bench_tile_ab.py --which real --rounds 5 > $OUT/${T}_real_tiles.log 2>> $OUT/${T}_bench.err
timeout 600 python tools/bench_tile_ab.py --which tile --rounds 5 > $OUT/${T}_tiles.log 2>> $OUT/${T}_bench.err
cut -c1-200 $OUT/${T}_real_tiles.log | grep -v "TH=\|MB=" ; grep -A3 "^2d c2c f32 128x128 K=8192\|^3d c2c f64 64^3 K=64" $OUT/${T}_tiles.log | cut -c1-160
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${T}_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-extra > $OUT/${T}_bench_under_ncu.log 2>&1
export BBFFT_CUDA_JIT_LINEINFO=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bbfft_.2.2d -c 6 -f -o $OUT/${T}_full_tiles \
    python tools/bench_tile_ab.py --which prof2 > $OUT/${T}_full_tiles.log 2>&1
ls -la $OUT | grep $T | cut -c1-120
