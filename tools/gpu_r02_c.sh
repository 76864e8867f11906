#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention" 2>&1 | tail -5
python tools/probe_attn.py > gpurun_out/r02_attn_probe.txt 2>&1
cat gpurun_out/r02_attn_probe.txt | tail -7
NJ=12 python tools/probe_attn_prof.py 0 > gpurun_out/r02_attn_prof.txt 2>&1
head -14 gpurun_out/r02_attn_prof.txt
