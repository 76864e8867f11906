#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -12 > gpurun_out/r02_pytest_gpu.log
cat gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02_bench.err > gpurun_out/r02_bench.json
tail -3 gpurun_out/r02_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks')}, d['e2e']['value'], d['roofline']['frac'])
for o in d['roofline']['others']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in o.items()})
for k,v in d['secondary'].items(): print(k, {kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','ms_per_batch','ms_encoder','ms_decode','ms_per_decode_step','error','clocks')})
PY
