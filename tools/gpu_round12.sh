#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_decode.py --ctc-weight 0.2 --steps 16 2>&1 | grep -v -i warn | head -16 | tee gpurun_out/profile_decode_joint.txt
timeout 300 python tools/profile_decode.py --batch 12 --beams 5 --ctc-weight 0.2 --steps 16 2>&1 | grep -v -i warn | head -18 | tee gpurun_out/profile_decode_beam.txt
