#!/bin/bash
# ncu evidence for profiles/: launch list of our kernels over one bench step + full captures of the top kernels
mkdir -p gpurun_out
K='regex:gemm_bf16_kernel|attention_fa_kernel|fddt_ln_kernel|fddt_ln_tma_kernel|features_to_cl_kernel|zero_pad_rows_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 229 -c 458 --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 24 -c 5 -f -o gpurun_out/prof_gemm_r01 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fa_kernel -s 4 -c 1 -f -o gpurun_out/prof_attn_r01 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fddt_ln_tma_kernel -s 6 -c 2 -f -o gpurun_out/prof_fddt_r01 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fddt.log 2>&1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_idle.csv
ls -la gpurun_out | tail -12
