#!/bin/bash
# full GPU parity suite (no -x: list every failure) + training / decode benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_train.py --workload finetune --steps 6 --warmup 3 2>gpurun_out/train.err | tee gpurun_out/train_finetune.json
timeout 600 python tools/bench_train.py --workload ctc_pretrain --steps 6 --warmup 3 2>>gpurun_out/train.err | tee gpurun_out/train_ctc.json
timeout 600 python tools/bench_decode.py 2>gpurun_out/decode.err | tee gpurun_out/decode.json
tail -5 gpurun_out/train.err gpurun_out/decode.err
