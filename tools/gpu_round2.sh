#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mel.py -m gpu -x -q -k "attention or mel" 2>&1 | tail -15 | tee gpurun_out/pytest_kernels.log
timeout 300 python tools/probe_attn.py 2>&1 | tee gpurun_out/probe_attn_v7.log
timeout 300 python tools/probe_mel.py 2>&1 | tee gpurun_out/probe_mel.log
