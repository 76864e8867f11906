#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_decoder.log
