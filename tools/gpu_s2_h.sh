#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --kernel-name "regex=attention_bwd_fused|colsum_bf16_vec" --error-exitcode 9 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py -m gpu -q -x -k "attention_backward or colsum" > gpurun_out/s2_sanitizer_c.log 2>&1
echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" gpurun_out/s2_sanitizer_c.log | head -8
