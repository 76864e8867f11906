#!/bin/bash
# One gpurun call: smoke, bench (N=1), ncu launch list + full captures of the top kernels.  Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 12 -c 4 -f -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 2 -f -o gpurun_out/prof_attn \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out
