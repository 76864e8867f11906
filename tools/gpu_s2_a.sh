#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ln_fddt_bwd_ring -s 100 -c 2 -f -o gpurun_out/prof_lnbwd_r02 \
    python tools/bench_train.py --workload finetune --steps 1 --warmup 3 > gpurun_out/ncu_lnbwd.log 2>&1
ls -la gpurun_out/prof_lnbwd_r02.ncu-rep; tail -2 gpurun_out/ncu_lnbwd.log | cut -c1-200
