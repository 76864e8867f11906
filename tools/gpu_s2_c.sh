#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py -q -x -k "gemm or gelu" 2>&1 | tail -3 | cut -c1-200
DICOW_GEMM_TAIL_SPLIT=0 ROWS=12000 timeout 300 python tools/bench_gemm_train.py 2>&1 | grep -A2 "^out \|^fc2 " | grep "form2\|^out\|^fc2" | cut -c1-160 | sed 's/^/nosplit /'
for v in 0 1 0 1; do
  DICOW_GEMM_TAIL_SPLIT=$v python tools/bench_train.py --workload finetune --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TAIL_SPLIT=$v finetune ms/step', round(d['ms_per_step'],2), d['clocks'])"
done
