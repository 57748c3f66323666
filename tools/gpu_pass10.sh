#!/bin/bash
# pass 10: validation of the round-1 state -- full GPU test tier, headline bench (+ reference arm),
# ncu launch list of the bench command, ncu --set full of the dominant / worst kernels, native
# benchmark against cuFFT
set -u
TAG=r01i
OUT=gpurun_out; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --per-size $OUT/${TAG}_per_size.csv > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>&1; cut -c1-200 $OUT/${TAG}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-extra > $OUT/${TAG}_bench_under_ncu.log 2>&1
for spec in 4:64 8:490 4:343 8:256; do
  fp=${spec%%:*}; n=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:bbfft -c 1 --launch-skip 3 -f -o $OUT/${TAG}_full_f${fp}_n${n} \
      python tools/sweep_gpu.py --fp $fp --sizes $n --check 0 > $OUT/${TAG}_full_f${fp}_n${n}.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bbfft_c2c2d -c 1 --launch-skip 2 -f -o $OUT/${TAG}_full_tile2d \
      python tools/bench_configs.py --which c4 > $OUT/${TAG}_full_tile2d.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bbfft_r2ch -c 1 --launch-skip 2 -f -o $OUT/${TAG}_full_r2c \
      python tools/bench_configs.py --which c3 > $OUT/${TAG}_full_r2c.log 2>&1
( timeout 600 tools/bin/bbfft-bench -o -m 16 sc 2 4 8 16 32 64 128 256 512 27 100 243 343 500; \
  timeout 600 tools/bin/bbfft-bench -o -m 16 dc 2 8 64 128 256 512 100 343 500; \
  timeout 300 tools/bin/bbfft-bench -o sr 256; timeout 300 tools/bin/bbfft-bench -i sr 256; \
  timeout 300 tools/bin/bbfft-bench -o -m 16 sr 64 256 500 ) > $OUT/${TAG}_native_bench_vs_cufft.csv 2>&1
tail -4 $OUT/${TAG}_native_bench_vs_cufft.csv
ls $OUT | grep $TAG
