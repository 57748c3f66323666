#!/bin/bash
# round 2, pass zf: the final tree's bench line on 2 GPUs, launched the way the driver launches it
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02zf_bench_2gpu.json 2> $OUT/r02zf_bench_2gpu.err
echo "bench rc=$?"; cut -c1-300 $OUT/r02zf_bench_2gpu.json; tail -3 $OUT/r02zf_bench_2gpu.err | cut -c1-300
