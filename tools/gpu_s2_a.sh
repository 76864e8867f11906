#!/bin/bash
for v in 0 1 0 1; do
DICOW_PREPARE_GRAPH=$v python tools/bench_train.py --workload ctc_pretrain --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PREPARE_GRAPH=$v ctc ms/step', round(d['ms_per_step'],2), d['clocks'])"
done
