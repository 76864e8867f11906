#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -x -k "lora" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_turbo_parity.py -q -k "finetune" 2>&1 | tail -3
K='regex:gemm_bf16|attention_fa_kernel|fddt_ln|features_to_cl|zero_pad_rows'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 229 -c 229 --csv \
    --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --secondary none --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
