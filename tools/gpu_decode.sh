#!/bin/bash
# One gpurun call for the decode path: parity tests, step benchmark (fused / unfused), SE-DiCoW end-to-end recipes
# (greedy, greedy + joint CTC, beam 5 + joint CTC), step timelines.  Outputs -> gpurun_out/
mkdir -p gpurun_out; rm -f gpurun_out/decode_modes.json gpurun_out/decode_recipes.json
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_ctc_joint.py tests/test_gpu_beam.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_decoder.log
for mode in "" "--ln-prologue" "--unfused"; do
  timeout 300 python tools/bench_decode.py $mode 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_modes.json
done
timeout 600 python tools/bench_decode.py --workload se_dicow 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json
timeout 600 python tools/bench_decode.py --workload se_dicow --ctc-weight 0.2 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json
timeout 600 python tools/bench_decode.py --workload se_dicow --batch 12 --beams 5 --ctc-weight 0.2 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_recipes.json
timeout 300 python tools/profile_decode.py 2>&1 | grep -v -i warn | tail -62 > gpurun_out/profile_decode_fused.txt
timeout 300 python tools/profile_decode.py --batch 12 --beams 5 --ctc-weight 0.2 --steps 16 2>&1 | grep -v -i warn | head -18 > gpurun_out/profile_decode_beam.txt
