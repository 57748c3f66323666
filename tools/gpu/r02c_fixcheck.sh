#!/bin/bash
# round 2, pass c: the c2r fix on the GPU -- repro soak, new parity shapes, callbacks, C++ device test
set -u
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02c_fixcheck.log
: > $L
R=tests/cpp/repro_c2r
export BBFFT_CUDA_KERNEL_CACHE=/tmp/kc; mkdir -p /tmp/kc
for st in blocking nonblocking; do
  timeout 300 $R c2r f64 32 424 4 300 fresh $st 0 2>&1 | tail -3 >> $L
  timeout 300 $R c2r f64 32 424 4 500 reuse $st 0 2>&1 | tail -3 >> $L
done
BBFFT_CUDA_KEEP_REGCAP=1 timeout 300 $R c2r f64 32 424 4 20 fresh blocking 0 2>&1 | tail -3 >> $L
for t in "ML=16" "ML=4" "BH=2" "ST=1" "T=53"; do
  BBFFT_CUDA_TUNE="$t" timeout 300 $R c2r f64 32 424 4 20 fresh blocking 0 2>&1 | tail -2 >> $L
done
timeout 300 $R c2r f64 16 424 4 50 fresh blocking 0 2>&1 | tail -2 >> $L
timeout 300 $R c2r f64 8 424 4 50 fresh blocking 0 2>&1 | tail -2 >> $L
timeout 300 $R c2c f64 16 509 8 50 fresh blocking 0 2>&1 | tail -2 >> $L
unset BBFFT_CUDA_KERNEL_CACHE
echo "== pytest new shapes + callbacks" >> $L
timeout 1500 python -m pytest tests/test_gpu_r2c.py tests/test_gpu_callback.py -x -q -m gpu 2>&1 | tail -8 >> $L
echo "== C++ device test" >> $L
( time timeout 1500 tests/cpp/test_cuda_api > $OUT/r02c_cpp_api.log 2>&1; echo "test_cuda_api rc=$?" ) >> $L 2>&1
tail -6 $OUT/r02c_cpp_api.log >> $L
echo "== sanitizers on the C++ callback section (synccheck on r2c single-stage kernels)" >> $L
timeout 300 compute-sanitizer --tool synccheck $R r2c f64 3 8 33 2 fresh blocking 0 2>&1 | tail -3 >> $L
timeout 300 compute-sanitizer --tool synccheck $R r2c f32 32 16 33 2 fresh blocking 0 2>&1 | tail -3 >> $L
cat $L | cut -c1-300
