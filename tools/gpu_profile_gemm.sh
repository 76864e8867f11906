#!/bin/bash
# launch list of one bench step (all of our kernels) + full capture of the GEMM family
mkdir -p gpurun_out
K='regex:gemm_bf16|attention_fa_kernel|fddt_ln|features_to_cl_kernel|zero_pad_rows_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 229 -c 458 --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 24 -c 5 -f -o gpurun_out/prof_gemm_r01 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
timeout 600 python tools/profile_train.py --workload finetune 2>&1 | grep -v -i warn | tail -40 > gpurun_out/profile_train_finetune.txt
timeout 600 python tools/bench_train.py --workload finetune --steps 6 --warmup 3 2>/dev/null | tee gpurun_out/train_finetune.json
ls -la gpurun_out | tail -6
