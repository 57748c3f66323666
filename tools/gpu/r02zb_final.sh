#!/bin/bash
# round 2, pass zb: final state with the X2 wisdom flags -- full GPU tier, bench line (+ per-size table), reference arm,
# r2c / c2r M=16 sweep, ncu launch list of the bench command, ncu --set full of the fused r2c tile kernel
set -u
OUT=gpurun_out; mkdir -p $OUT
T=r02zb
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/${T}_smi.txt 2>&1
( time timeout 1700 python -m pytest tests -m gpu -x -q ) > $OUT/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${T}_pytest.log
tail -6 $OUT/${T}_pytest.log
timeout 1500 python bench.py --per-size $OUT/${T}_per_size.csv > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench rc=$?"
cut -c1-300 $OUT/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > $OUT/${T}_bench_reference.json 2>> $OUT/${T}_bench.err; echo "reference rc=$?"
export BBFFT_CUDA_KERNEL_CACHE=$PWD/kcache BBFFT_CUDA_JIT_LINEINFO=0
timeout 600 python tools/bench_configs.py --which none --real-sweep > $OUT/${T}_real_sweep.jsonl 2>> $OUT/${T}_bench.err
python - <<PY
import json
rows=[json.loads(l) for l in open("$OUT/${T}_real_sweep.jsonl") if l.startswith("{")]
peak=6534.5
fr=sorted((r["GBs"]/peak, r["config"], r["fp"], r["shape"]) for r in rows)
print("real sweep: rows=%d min=%.3f median=%.3f n<0.8=%d  lowest: %s" % (len(fr), fr[0][0], fr[len(fr)//2][0], sum(1 for f in fr if f[0]<0.8), [(round(f[0],3),f[1],f[2],f[3]) for f in fr[:6]]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${T}_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-extra > $OUT/${T}_bench_under_ncu.log 2>&1
export BBFFT_CUDA_JIT_LINEINFO=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bbfft_r2c2d -c 1 --launch-skip 2 -f -o $OUT/${T}_full_r2c_tile \
    python tools/bench_tile_ab.py --which prof2 > $OUT/${T}_full_tiles.log 2>&1
ls -la $OUT | grep $T | cut -c1-120
