#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_training.py -q -x -k "adamw or updates_and_second" 2>&1 | tail -3 | cut -c1-200
for v in 1 0 1 0; do
  DICOW_TORCH_ADAMW=$v python tools/bench_train.py --workload finetune --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TORCH_ADAMW=$v finetune ms/step', round(d['ms_per_step'],2), d['clocks'], d['optimizer'][:30])"
done
python tools/profile_train.py --workload finetune 2>&1 | grep -E "adamw|Adam|multi_tensor|device span|total GPU" | cut -c1-150
