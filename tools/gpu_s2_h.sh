#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --kernel-name "regex=decode_cross_attention_ring|adamw_kernel" --error-exitcode 9 \
    python -m pytest tests/test_gpu_decoder.py tests/test_gpu_training.py -m gpu -q -x -k "greedy_decode_matches or adamw or speculative" > gpurun_out/s2_sanitizer_d.log 2>&1
echo "rc=$?"; grep -E "passed|failed|ERROR SUMMARY|Invalid" gpurun_out/s2_sanitizer_d.log | head -5
timeout 900 compute-sanitizer --tool racecheck --kernel-name "regex=decode_cross_attention_ring" --error-exitcode 9 \
    python -m pytest tests/test_gpu_decoder.py -m gpu -q -x -k "greedy_decode_matches" > gpurun_out/s2_sanitizer_e.log 2>&1
echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/s2_sanitizer_e.log | head -5
