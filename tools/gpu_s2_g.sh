#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/s2_longform.txt
timeout 600 python tools/bench_longform.py --batch 16 --windows 8 --tokens 64 2>&1 | tail -2 | tee -a gpurun_out/s2_longform.txt | cut -c130-300
timeout 600 python tools/bench_longform.py --batch 16 --windows 8 --tokens 128 --se 2>&1 | tail -2 | tee -a gpurun_out/s2_longform.txt | cut -c130-300
timeout 600 python tools/bench_longform.py --batch 4 --windows 8 --tokens 128 2>&1 | tail -2 | tee -a gpurun_out/s2_longform.txt | cut -c130-300
timeout 600 python tools/bench_longform.py --batch 16 --windows 4 --tokens 64 --sms 64 2>&1 | tail -1 | tee -a gpurun_out/s2_longform.txt | cut -c130-300
timeout 600 python tools/bench_longform.py --batch 16 --windows 4 --tokens 64 --sms 96 2>&1 | tail -1 | tee -a gpurun_out/s2_longform.txt | cut -c130-300
t0=$(date +%s)
timeout 1200 python bench.py 2>gpurun_out/s2_bench.err | tee gpurun_out/s2_bench.json | cut -c1-100
echo "bench wall: $(( $(date +%s) - t0 )) s"
