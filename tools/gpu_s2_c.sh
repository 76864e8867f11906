#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s2_full.log 2>&1; echo "full rc=$?"; tail -3 gpurun_out/s2_full.log | cut -c1-300
python tools/bench_train.py --workload ctc_pretrain --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctc ms/step', round(d['ms_per_step'],2), d['clocks'])"
python tools/bench_train.py --workload finetune --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('finetune ms/step', round(d['ms_per_step'],2), d['clocks'])"
