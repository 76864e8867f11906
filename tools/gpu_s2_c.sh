#!/bin/bash
# A/B of two library builds: the in-tree .so is swapped (the box copy is scratch)
L=ts-asr-whisper_b200/libdicow_b200.so
for v in nohint hint nohint hint; do
  cp tools/bin/libdicow_$v.so $L
  python tools/bench_train.py --workload finetune --steps 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v finetune ms/step', round(d['ms_per_step'],2), d['clocks'])"
  python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v headline', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], [(o['kernel'], round(o['ms_per_step'],2)) for o in d['roofline']['others']])"
done
cp tools/bin/libdicow_hint.so $L
DICOW_ATTN_BWD_FUSED=1 timeout 120 python tools/bench_attn_bwd.py 2>&1 | tail -1 | sed 's/^/hint /'
cp tools/bin/libdicow_nohint.so $L
DICOW_ATTN_BWD_FUSED=1 timeout 120 python tools/bench_attn_bwd.py 2>&1 | tail -1 | sed 's/^/nohint /'
