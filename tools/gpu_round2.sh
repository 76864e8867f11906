#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "attention_backward" 2>&1 | tail -25 | tee gpurun_out/pytest_attn_bwd.log
