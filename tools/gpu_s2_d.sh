#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention_backward" 2>&1 | tail -3 | cut -c1-250
for d in 0 1 2 3; do DICOW_BWD_FUSED_DBG=$d timeout 120 python tools/bench_attn_bwd.py 2>&1 | tail -1 | sed "s/^/dbg=$d /"; done
