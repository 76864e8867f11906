#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm" 2>&1 | tail -15 | tee gpurun_out/pytest_gemm.log
