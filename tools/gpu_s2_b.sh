#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py -q -x -k "gemm or gelu" 2>&1 | tail -15
timeout 300 python tools/bench_gemm_train.py 2>&1 | tee gpurun_out/s2_bench_gemm_train.txt
