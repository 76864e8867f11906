#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python bench.py 2>gpurun_out/s2_bench.err | tee gpurun_out/s2_bench.json | cut -c1-300
echo "bench wall: $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/s2_bench.err
