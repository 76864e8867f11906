#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_decode.py 2>&1 | grep -v Warning | tail -70 | tee gpurun_out/profile_decode_fused.txt
DICOW_DISABLE_PDL=1 timeout 300 python tools/profile_decode.py 2>&1 | grep -v Warning | head -14 | tee gpurun_out/profile_decode_fused_nopdl.txt
timeout 300 python tools/profile_decode.py --unfused 2>&1 | grep -v Warning | head -14 | tee gpurun_out/profile_decode_unfused.txt
