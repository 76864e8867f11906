#!/bin/bash
# round 2, call A: new turbo-dimension parity tests, full GPU suite, bench with the secondary workloads
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_turbo_parity.py -q -s --durations=10 2>&1 | tail -60 > gpurun_out/r02_turbo_parity.log
tail -25 gpurun_out/r02_turbo_parity.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_turbo_parity.py --durations=8 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu.log
tail -15 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/r02_bench.err > gpurun_out/r02_bench.json
tail -3 gpurun_out/r02_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks')}, d['e2e']['value'], d['roofline']['frac'])
for o in d['roofline']['others']: print(o)
for k,v in d['secondary'].items(): print(k, json.dumps(v)[:900])
print(d.get('cpu_baseline'))
PY
