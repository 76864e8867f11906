#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_kernels.py -m gpu -x -q -k "encoder or golden or fddt" 2>&1 | tail -5 | tee gpurun_out/pytest_enc.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_v8.json
