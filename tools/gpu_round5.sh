#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/pytest_decoder.log
for mode in "" "--unfused"; do
  timeout 300 python tools/bench_decode.py $mode 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_modes.json
done
DICOW_DISABLE_PDL=1 timeout 300 python tools/bench_decode.py 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_modes.json
timeout 300 python tools/bench_decode.py --no-graphs 2>>gpurun_out/decode.err | tee -a gpurun_out/decode_modes.json
timeout 600 python tools/bench_decode.py --workload se_dicow 2>>gpurun_out/decode.err | tee gpurun_out/decode_se_dicow.json
K='regex:gemm_skinny|decode_attention|embed_kernel|logits_rules|fddt_ln|advance_kernel|decode_linear'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 200 -c 108 --csv \
    --log-file gpurun_out/launches_decode_fused.csv python tools/bench_decode.py --steps 8 --no-graphs > gpurun_out/decode_under_ncu.log 2>&1
tail -3 gpurun_out/decode.err
