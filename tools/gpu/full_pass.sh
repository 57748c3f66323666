#!/bin/bash
# One GPU-box pass: parity tests, bench (+ per-size table), ncu launch list of the bench command,
# ncu --set full captures of selected kernels.  Outputs under gpurun_out/.
set -u
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --per-size $OUT/${TAG}_per_size.csv > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"
cat $OUT/${TAG}_bench.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
for spec in 4:64 4:512 8:350 8:512; do
  fp=${spec%%:*}; n=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bbfft -c 1 --launch-skip 3 -f -o $OUT/${TAG}_full_f${fp}_n${n} \
      python tools/sweep_gpu.py --fp $fp --sizes $n --check 0 > $OUT/${TAG}_full_f${fp}_n${n}.log 2>&1
done
ls -la $OUT | tail -30
