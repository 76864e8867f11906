#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decoder.py -q -x 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_turbo_parity.py -q -x -k "greedy or follows" 2>&1 | tail -3
python tools/probe_decode_mega.py 16 2>&1 | tee gpurun_out/r02_decode_mega_phases.txt | tail -22
for m in 1 0; do DICOW_DECODE_MEGA=$m timeout 300 python tools/bench_decode.py --batch 16 --steps 128 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mega=$m', d['ms_per_step'], d['frac_of_hbm_roofline'], d['generated_tail'])"; done
