#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_gpu_beam.py tests/test_gpu_decoder.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python tools/profile_decode.py --batch 12 --beams 5 --ctc-weight 0.2 --steps 16 2>&1 | grep -v -i warn | head -9 | tee gpurun_out/profile_decode_beam.txt
DICOW_MQ_ATTENTION=0 timeout 300 python tools/profile_decode.py --batch 12 --beams 5 --ctc-weight 0.2 --steps 16 2>&1 | grep -v -i warn | head -6
timeout 600 python tools/bench_decode.py --workload se_dicow --batch 12 --beams 5 --ctc-weight 0.2 2>>gpurun_out/decode.err | cut -c1-420
