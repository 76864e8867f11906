#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 300 python tools/bench_decode.py 2>&1 | tail -2 | tee gpurun_out/bench_decode.json
timeout 300 python tools/bench_decode.py --no-graphs 2>&1 | tail -1 | tee gpurun_out/bench_decode_nograph.json
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_v5.json
