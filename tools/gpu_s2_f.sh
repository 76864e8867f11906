#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_decoder.py -q -x -k "speculative or se_dicow_matches" 2>&1 | tail -1 | cut -c1-200
echo "--- DiCoW 8 windows, 64 tokens, auto"
timeout 600 python tools/bench_longform.py --batch 16 --windows 8 --tokens 64 2>&1 | tail -2 | cut -c130-500
echo "--- SE-DiCoW 8 windows, 128 tokens, auto"
timeout 600 python tools/bench_longform.py --batch 16 --windows 8 --tokens 128 --se 2>&1 | tail -2 | cut -c130-500
echo "--- DiCoW 4 recordings, 8 windows, 128 tokens, auto"
timeout 600 python tools/bench_longform.py --batch 4 --windows 8 --tokens 128 2>&1 | tail -2 | cut -c130-500
