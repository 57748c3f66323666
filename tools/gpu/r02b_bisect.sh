#!/bin/bash
# round 2, pass b: bisect the wrong c2r result (fp64 M=32 N=424) over planner overrides and shapes
set -u
OUT=gpurun_out; mkdir -p $OUT
R=tests/cpp/repro_c2r
L=$OUT/r02b_bisect.log
: > $L
export BBFFT_CUDA_KERNEL_CACHE=/tmp/kc; mkdir -p /tmp/kc
run() { echo "### TUNE='${BBFFT_CUDA_TUNE:-}' $*" >> $L; timeout 120 $R "$@" 2>&1 | grep -v "^  iter [1-9]" | cut -c1-1500 >> $L; }
run c2r f64 32 424 4 3 fresh blocking 0
for t in "MB=1" "T=53" "T=32" "T=16" "RF=0" "ML=16" "ML=4" "R=2x2x53" "BH=2" "LD=1" "ST=1"; do
  BBFFT_CUDA_TUNE="$t" run c2r f64 32 424 4 3 fresh blocking 0
done
run c2r f32 32 424 4 3
run c2r f64 8 424 4 3
run c2r f64 16 424 4 3
run c2r f64 32 212 4 3
run c2r f64 32 318 4 3
run c2r f64 32 530 4 3
run c2r f64 32 106 4 3
run c2r f64 32 428 4 3
run c2r f64 32 148 4 3
run c2r f64 1 424 33 3
run c2r f64 3 424 33 3
run r2c f64 32 424 4 3
run r2c f64 32 212 4 3
run r2c f32 32 424 4 3
run c2c f64 32 212 4 3
run c2c f64 32 106 4 3
run c2c f32 16 212 8 3
run c2r f64 32 128 4 3
run c2r f64 32 420 4 3
run c2r f64 32 490 4 3
cat $L | cut -c1-400
