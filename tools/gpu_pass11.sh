#!/bin/bash
# pass 11 (2 GPUs): the bench launched the way the driver does for N > 1, and the native multi-device benchmark
set -u
TAG=r01j
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/${TAG}_bench_2gpu.json 2> $OUT/${TAG}_bench_2gpu.err
cut -c1-300 $OUT/${TAG}_bench_2gpu.json; tail -3 $OUT/${TAG}_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $OUT/${TAG}_bench_2gpu_reference.json 2>&1
cut -c1-200 $OUT/${TAG}_bench_2gpu_reference.json
( timeout 300 tools/bin/bbfft-bench -o -m 16 --impl bbfft --devices 1 sc 64 256 500; timeout 300 tools/bin/bbfft-bench -o -m 16 --impl bbfft --devices 2 -k 262144 sc 64 ) > $OUT/${TAG}_native_devices.csv 2>&1
cat $OUT/${TAG}_native_devices.csv | tail -8
ls $OUT | grep $TAG
