#!/bin/bash
# round 2, pass j: the measurements of record -- bench line (+ per-size table), reference arm, ncu launch list of
# the bench command, ncu --set full captures, the other BASELINE configs, the real sweep, native bench, tft example
set -u
OUT=gpurun_out; mkdir -p $OUT
T=${1:-r02j}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/${T}_smi.txt 2>&1
timeout 1500 python bench.py --per-size $OUT/${T}_per_size.csv > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"
cut -c1-400 $OUT/${T}_bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 3 > $OUT/${T}_bench_reference.json 2>> $OUT/${T}_bench.err; echo "reference rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${T}_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-extra > $OUT/${T}_bench_under_ncu.log 2>&1
for spec in 4:64 8:490 8:343 4:343; do
  fp=${spec%%:*}; n=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:bbfft -c 1 --launch-skip 3 -f -o $OUT/${T}_full_f${fp}_n${n} \
      python tools/sweep_gpu.py --fp $fp --sizes $n --check 0 > $OUT/${T}_full_f${fp}_n${n}.log 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bbfft_c2c2d -c 1 --launch-skip 3 -f -o $OUT/${T}_full_tile2d \
    python tools/bench_configs.py --which c4 > $OUT/${T}_full_tile2d.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:bbfft_r2ch -c 1 --launch-skip 3 -f -o $OUT/${T}_full_c3_r2c \
    python tools/bench_configs.py --which c3 > $OUT/${T}_full_c3_r2c.log 2>&1
timeout 900 python tools/bench_configs.py --which c1,c3,c4,c5 > $OUT/${T}_configs.jsonl 2>> $OUT/${T}_bench.err
timeout 900 python tools/bench_configs.py --which none --real-sweep > $OUT/${T}_real_sweep.jsonl 2>> $OUT/${T}_bench.err
timeout 300 tools/bin/bbfft-bench -o -m 1 -k 16384 --burst 200 --impl both sc 64 > $OUT/${T}_native.csv 2>&1
timeout 600 tools/bin/bbfft-bench -o -m 16 --impl both sc 64 256 490 512 >> $OUT/${T}_native.csv 2>&1
timeout 600 tools/bin/bbfft-bench -o -m 16 --impl both dc 64 256 490 512 >> $OUT/${T}_native.csv 2>&1
timeout 600 examples/bin/tft-cuda > $OUT/${T}_tft.txt 2>&1
ls -la $OUT | grep $T | cut -c1-120
