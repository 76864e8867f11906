#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/decode_recipes.json
timeout 600 python tools/bench_decode.py --workload se_dicow 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json | cut -c1-600
timeout 600 python tools/bench_decode.py --workload se_dicow --ctc-weight 0.2 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json | cut -c1-600
timeout 600 python tools/bench_decode.py --workload se_dicow --batch 12 --beams 5 --ctc-weight 0.2 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json | cut -c1-600
timeout 600 python tools/bench_decode.py --workload se_dicow --batch 12 --beams 5 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json | cut -c1-600
tail -3 gpurun_out/decode.err
