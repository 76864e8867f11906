#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_longform.py --batch 16 --windows 4 --tokens 64 2>&1 | tail -2 | tee gpurun_out/s2_longform.txt
timeout 600 python tools/bench_longform.py --batch 16 --windows 4 --tokens 128 --se 2>&1 | tail -2 | tee -a gpurun_out/s2_longform.txt
timeout 600 python tools/bench_longform.py --batch 4 --windows 4 --tokens 128 2>&1 | tail -2 | tee -a gpurun_out/s2_longform.txt
timeout 600 python tools/bench_longform.py --batch 16 --windows 4 --tokens 128 --se --sms 64 2>&1 | tail -2 | tee -a gpurun_out/s2_longform.txt
