#!/bin/bash
mkdir -p gpurun_out
NJ=12 python tools/probe_attn_prof.py 0 4 > gpurun_out/r02_attn_prof.txt 2>&1
python tools/probe_attn.py > gpurun_out/r02_attn_probe.txt 2>&1
tail -5 gpurun_out/r02_attn_probe.txt
timeout 900 python -m pytest tests/test_gpu_turbo_parity.py -q -s -k finetune 2>&1 | grep -E "turbo|passed|failed" > gpurun_out/r02_turbo_ft.log
cat gpurun_out/r02_turbo_ft.log
