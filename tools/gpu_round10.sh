#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py tests/test_gpu_backward.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gemm.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-1700
tail -3 gpurun_out/bench.err
timeout 600 python tools/bench_train.py --workload finetune --steps 6 --warmup 3 2>/dev/null | tee gpurun_out/train_finetune.json | cut -c1-300
