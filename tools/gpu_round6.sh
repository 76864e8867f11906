#!/bin/bash
mkdir -p gpurun_out
K='regex:decode_attention|decode_linear'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 36 -c 9 -f -o gpurun_out/prof_decode \
    python tools/bench_decode.py --steps 8 --no-graphs > gpurun_out/ncu_decode.log 2>&1
tail -2 gpurun_out/ncu_decode.log
ls -la gpurun_out/*.ncu-rep
