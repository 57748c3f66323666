#!/bin/bash
# round 2, pass g: override searches -- the sweep sizes below 0.8, BASELINE config 3 (M = 1 real N = 256) and
# config 1 (graph replay); candidates were compiled on the build box (kcache/)
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02g_tune.log
: > $L
timeout 1500 python tools/tune_list.py --cases tools/cases_laggards.json --out $OUT/r02g_laggards.json > $OUT/r02g_laggards.log 2>&1
grep -c rejected $OUT/r02g_laggards.log >> $L
grep -v rejected $OUT/r02g_laggards.log | awk '{c[$1]++; if (c[$1]<=4) print}' >> $L
timeout 1500 python tools/tune_list.py --cases tools/cases_c3.json --out $OUT/r02g_c3.json > $OUT/r02g_c3.log 2>&1
grep -v rejected $OUT/r02g_c3.log | awk '{c[$1]++; if (c[$1]<=5) print}' >> $L
timeout 900 python tools/tune_list.py --cases tools/cases_c1.json --graph --inner 20 --filler 0 --out $OUT/r02g_c1.json > $OUT/r02g_c1.log 2>&1
grep -v rejected $OUT/r02g_c1.log | awk '{c[$1]++; if (c[$1]<=6) print}' >> $L
cat $L | cut -c1-250
