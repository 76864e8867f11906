#!/bin/bash
# round 2 ncu evidence for profiles/: launch list of one encoder step + --set full captures of the dominant kernels
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --secondary none --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 229 -c 229 --csv \
    --log-file gpurun_out/launches_r02.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2cta_kernel -s 24 -c 5 -f -o gpurun_out/prof_gemm_r02 \
    $BENCH > gpurun_out/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fa_kernel -s 4 -c 1 -f -o gpurun_out/prof_attn_r02 \
    $BENCH > gpurun_out/ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fddt_ln_tma_kernel -s 6 -c 2 -f -o gpurun_out/prof_fddt_r02 \
    $BENCH > gpurun_out/ncu_fddt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel -s 6 -c 3 -f -o gpurun_out/prof_mel_r02 \
    python tools/probe_mel.py > gpurun_out/ncu_mel.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_attention_kernel -s 40 -c 2 -f -o gpurun_out/prof_decattn_r02 \
    python tools/bench_decode.py --steps 8 --no-graphs > gpurun_out/ncu_decattn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_kernel -s 4 -c 2 -f -o gpurun_out/prof_attnbwd_r02 \
    python tools/bench_train.py --workload finetune --steps 1 --warmup 1 > gpurun_out/ncu_attnbwd.log 2>&1
python tools/probe_mel.py > gpurun_out/r02_mel_probe.txt 2>&1
ls -la gpurun_out | tail -14
