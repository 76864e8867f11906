#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_backward.py tests/test_gpu_encoder.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/pytest_se.log
timeout 600 python tools/bench_decode.py --workload se_dicow 2>gpurun_out/se.err | tee gpurun_out/decode_se_dicow.json
tail -3 gpurun_out/se.err
timeout 600 python tools/profile_train.py --workload finetune 2>&1 | tail -50 | tee gpurun_out/profile_train_finetune.txt
K='regex:gemm_skinny|decode_attention|embed_kernel|logits_rules|fddt_ln|advance_kernel'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 300 -c 171 --csv \
    --log-file gpurun_out/launches_decode.csv python tools/bench_decode.py --steps 8 --no-graphs > gpurun_out/decode_under_ncu.log 2>&1
tail -2 gpurun_out/decode_under_ncu.log
