#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "attention_backward" 2>&1 | tail -2 | cut -c1-250
for d in 0 1 2; do DICOW_BWD_FUSED_DBG=$d timeout 120 python tools/bench_attn_bwd.py 2>&1 | tail -1 | sed "s/^/dbg=$d /"; done
for v in 1 2; do
python tools/bench_train.py --workload finetune --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('finetune ms/step', round(d['ms_per_step'],2), d['clocks'])"
done
