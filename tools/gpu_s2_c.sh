#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_turbo_parity.py -q -x -k "not megakernel" 2>&1 | tail -3 | cut -c1-200
for v in 1 2 3; do
  python tools/bench_train.py --workload finetune --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('finetune ms/step', round(d['ms_per_step'],2), d['clocks'], round(d['peak_mem_gb'],1))"
done
python tools/profile_train.py --workload finetune 2>&1 | head -6 | cut -c1-250
