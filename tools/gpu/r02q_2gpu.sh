#!/bin/bash
# round 2, pass q: the bench line on 2 GPUs, launched the way the driver launches it, and the reference arm under torchrun
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02q_bench_2gpu.json 2> $OUT/r02q_bench_2gpu.err
echo "bench rc=$?"; cut -c1-300 $OUT/r02q_bench_2gpu.json; tail -3 $OUT/r02q_bench_2gpu.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > $OUT/r02q_bench_2gpu_reference.json 2>> $OUT/r02q_bench_2gpu.err
echo "reference rc=$?"; cut -c1-200 $OUT/r02q_bench_2gpu_reference.json
timeout 600 python -m pytest tests/test_gpu_z_cpp_api.py -x -q -m gpu -k soak 2>&1 | tail -3
