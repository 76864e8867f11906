#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/decode_recipes.json
timeout 800 python -m pytest tests/test_gpu_beam.py tests/test_gpu_ctc_joint.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python tools/bench_decode.py --workload se_dicow 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json | cut -c1-420
timeout 600 python tools/bench_decode.py --workload se_dicow --ctc-weight 0.2 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json | cut -c1-420
timeout 600 python tools/bench_decode.py --workload se_dicow --batch 12 --beams 5 --ctc-weight 0.2 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json | cut -c1-420
timeout 300 python tools/profile_decode.py --batch 12 --beams 5 --ctc-weight 0.2 --steps 16 2>&1 | grep -v -i warn | head -16 | tee gpurun_out/profile_decode_beam.txt
