#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_decoder.log
rm -f gpurun_out/decode_modes.json
for mode in "" "--ln-prologue" "--unfused"; do
  timeout 300 python tools/bench_decode.py $mode 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_modes.json
done
timeout 300 python tools/profile_decode.py 2>&1 | grep -v -i warn | tail -60 | tee gpurun_out/profile_decode_fused.txt
timeout 600 python tools/bench_decode.py --workload se_dicow 2>>gpurun_out/decode.err | tee gpurun_out/decode_se_dicow.json
