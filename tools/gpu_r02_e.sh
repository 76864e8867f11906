#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -q -x 2>&1 | tail -5
python tools/energy_probe.py --only gemm 2>&1 | grep -v "cuBLAS\|single-CTA" | tee gpurun_out/r02_energy_gemm_tma.txt
DICOW_GEMM_TMA_STORE=0 python tools/energy_probe.py --only gemm 2>&1 | grep -v "cuBLAS\|single-CTA" | tee gpurun_out/r02_energy_gemm_notma.txt
timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TMA store:', d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'])"
DICOW_GEMM_TMA_STORE=0 timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('st.global:', d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'])"
timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TMA store:', d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'])"
DICOW_GEMM_TMA_STORE=0 timeout 600 python bench.py --steps 10 --warmup 3 --secondary none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('st.global:', d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'])"
