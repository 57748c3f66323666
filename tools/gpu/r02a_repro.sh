#!/bin/bash
# round 2, pass a: chase the nondeterministic c2r result (fp64 M=32 N=424) on the GPU
set -u
OUT=gpurun_out; mkdir -p $OUT
R=tests/cpp/repro_c2r
L=$OUT/r02a_repro.log
: > $L
nvidia-smi --query-gpu=name,driver_version --format=csv >> $L
for st in blocking nonblocking default; do
  timeout 300 $R c2r f64 32 424 4 30 fresh $st 0 >> $L 2>&1
done
export BBFFT_CUDA_KERNEL_CACHE=/tmp/kc; mkdir -p /tmp/kc
for st in blocking nonblocking default; do
  timeout 300 $R c2r f64 32 424 4 400 fresh $st 0 >> $L 2>&1
  timeout 300 $R c2r f64 32 424 4 400 fresh $st 1 >> $L 2>&1
  timeout 300 $R c2r f64 32 424 4 400 reuse $st 0 >> $L 2>&1
done
timeout 300 $R c2r f32 32 424 4 400 fresh nonblocking 0 >> $L 2>&1
timeout 300 $R r2c f64 32 848 4 400 fresh nonblocking 0 >> $L 2>&1
timeout 300 $R c2c f64 16 490 64 400 fresh nonblocking 0 >> $L 2>&1
for tool in memcheck racecheck initcheck synccheck; do
  echo "== compute-sanitizer $tool" >> $L
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 $R c2r f64 32 424 4 3 fresh blocking 0 > $OUT/r02a_sanitizer_$tool.log 2>&1
  tail -8 $OUT/r02a_sanitizer_$tool.log >> $L
done
echo "== full C++ device test" >> $L
unset BBFFT_CUDA_KERNEL_CACHE
timeout 1500 tests/cpp/test_cuda_api > $OUT/r02a_cpp_api.log 2>&1; echo "test_cuda_api rc=$?" >> $L
tail -12 $OUT/r02a_cpp_api.log >> $L
cat $L | cut -c1-200
