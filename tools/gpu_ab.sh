#!/bin/bash
# A/B on ONE box, interleaved: pipelined vs serial accumulator drain of the GEMM epilogue
mkdir -p gpurun_out; rm -f gpurun_out/ab.json
for i in 1 2 3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'drain':'pipelined','value':d['value'],'sm_mhz':d['clocks']['sm_mhz'],'gemm_ms':d['kernels']['gemm']['ms_per_step'],'attn_ms':d['kernels']['attention']['ms_per_step']}))" | tee -a gpurun_out/ab.json
  DICOW_GEMM_SERIAL_DRAIN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'drain':'serial','value':d['value'],'sm_mhz':d['clocks']['sm_mhz'],'gemm_ms':d['kernels']['gemm']['ms_per_step'],'attn_ms':d['kernels']['attention']['ms_per_step']}))" | tee -a gpurun_out/ab.json
done
