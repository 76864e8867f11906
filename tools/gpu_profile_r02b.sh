#!/bin/bash
# round 2 (second session) ncu evidence: launch list of one fine-tune step, --set full of the single-pass attention backward and of the
# CTA-pair GEMM with MN-major operands (dgrad / wgrad / GELU-save)
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 3300 -c 1700 --csv \
    --log-file gpurun_out/launches_train_r02.csv python tools/profile_train.py --workload finetune > gpurun_out/ncu_launches_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_fused_kernel -s 3 -c 1 -f \
    -o gpurun_out/prof_attnbwd_fused_r02 python tools/bench_attn_bwd.py > gpurun_out/ncu_attnbwd_fused.log 2>&1
ROWS=12000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2cta_kernel -s 318 -c 5 -f \
    -o gpurun_out/prof_gemm_train_r02 python tools/bench_gemm_train.py > gpurun_out/ncu_gemm_train.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_train_r02.csv | tail -6
tail -3 gpurun_out/ncu_gemm_train.log
