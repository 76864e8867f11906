#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_training.py -q -x -k "cast_bf16_padded or ctc_pretrain" 2>&1 | tail -2 | cut -c1-200
python tools/profile_train.py --workload ctc_pretrain 2>&1 | grep -E "cast_2d|total GPU" | cut -c1-120
python tools/bench_train.py --workload ctc_pretrain --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctc ms/step', round(d['ms_per_step'],2), d['clocks'])"
