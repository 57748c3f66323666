#!/bin/bash
set -u
TAG=r01e
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --per-size $OUT/${TAG}_per_size.csv > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_bench.json
timeout 900 python tools/bench_configs.py --which c3 --real-sweep > $OUT/${TAG}_real_sweep.jsonl 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-extra > $OUT/${TAG}_bench_under_ncu.log 2>&1
for spec in 4:64 8:441 4:500; do
  fp=${spec%%:*}; n=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bbfft -c 1 --launch-skip 3 -f -o $OUT/${TAG}_full_f${fp}_n${n} \
      python tools/sweep_gpu.py --fp $fp --sizes $n --check 0 > $OUT/${TAG}_full_f${fp}_n${n}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bbfft_c2c2d -c 2 --launch-skip 2 -f -o $OUT/${TAG}_full_tile2d \
      python tools/bench_configs.py --which c4 > $OUT/${TAG}_full_tile2d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bbfft_c2rh -c 1 --launch-skip 2 -f -o $OUT/${TAG}_full_c2r \
      python tools/bench_configs.py --which c3 > $OUT/${TAG}_full_c2r.log 2>&1
ls $OUT | grep $TAG
