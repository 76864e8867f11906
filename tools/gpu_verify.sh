#!/bin/bash
# One gpurun call: full GPU parity suite, smoke, bench (N=1, with the secondary workloads) and the reference arm.  Outputs -> gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-600
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_ref.json | cut -c1-400
