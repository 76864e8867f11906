#!/bin/bash
echo "--- split"; python tools/probe_decode_attn.py 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_beam.py tests/test_gpu_ctc_joint.py tests/test_gpu_turbo_parity.py -q -x -k "not megakernel" 2>&1 | tail -2 | cut -c1-200
for v in 1 2; do python tools/bench_decode.py --workload se_dicow 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split', round(d['ms_per_batch'],2), round(d['ms_per_decode_step'],4), d.get('clocks'))"; done
