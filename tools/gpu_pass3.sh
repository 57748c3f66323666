#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/p3_pytest.log 2>&1; tail -3 $OUT/p3_pytest.log
timeout 600 python tools/bench_configs.py --which c3,c4 > $OUT/p3_configs.jsonl 2> $OUT/p3_configs.err; tail -2 $OUT/p3_configs.err
for bb in 0 8388608 16777216 33554432 50331648 67108864; do
  echo "== ND block bytes $bb"; BBFFT_CUDA_ND_BLOCK_BYTES=$bb timeout 300 python tools/bench_configs.py --which c4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print('   %-10s %-18s %8.1f us %6.0f GB/s L%d'%(r['config'],r['shape'],r['time_us'],r['GBs'],r['launches']))"
done > $OUT/p3_nd_block.txt 2>&1
echo "== multipass (no fusion)" >> $OUT/p3_nd_block.txt
for bb in 0 25165824; do BBFFT_CUDA_ND_FUSE=0 BBFFT_CUDA_ND_BLOCK_BYTES=$bb timeout 300 python tools/bench_configs.py --which c4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print('   %-10s %-18s %8.1f us %6.0f GB/s L%d'%(r['config'],r['shape'],r['time_us'],r['GBs'],r['launches']))"; done >> $OUT/p3_nd_block.txt 2>&1
cat $OUT/p3_nd_block.txt
timeout 600 python tools/exp_real_m1.py > $OUT/p3_exp_m1.txt 2>&1; tail -40 $OUT/p3_exp_m1.txt
timeout 900 python tools/bench_configs.py --which none --real-sweep > $OUT/p3_real_sweep.jsonl 2>&1
timeout 600 python bench.py --per-size $OUT/p3_per_size.csv --e2e-steps 0 --no-cpu-baseline > $OUT/p3_bench.json 2> $OUT/p3_bench.err; cut -c1-400 $OUT/p3_bench.json
timeout 300 tools/bin/bbfft-bench -o -m 16 sc 64 256 500 > $OUT/p3_native.csv 2>&1; timeout 300 tools/bin/bbfft-bench -o sr 256 >> $OUT/p3_native.csv 2>&1; timeout 300 tools/bin/bbfft-bench -i -v sr 256 >> $OUT/p3_native.csv 2>&1; cat $OUT/p3_native.csv
