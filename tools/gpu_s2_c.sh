#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do
  DICOW_ATTN_BWD_FUSED=$v python tools/bench_train.py --workload finetune --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FUSED_BWD=$v finetune ms/step', round(d['ms_per_step'],2), d['clocks'])"
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s2_full.log 2>&1; echo "full rc=$?"; tail -3 gpurun_out/s2_full.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_turbo_parity.py -q -x > gpurun_out/s2_c.log 2>&1; echo "pair rc=$?"; tail -1 gpurun_out/s2_c.log | cut -c1-200
