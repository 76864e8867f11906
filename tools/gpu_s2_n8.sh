#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/s2_bench_n8.err | tee gpurun_out/s2_bench_n8.json | cut -c1-200
tail -3 gpurun_out/s2_bench_n8.err
