mkdir -p gpurun_out
DICOW_PDL=1 timeout 600 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_beam.py tests/test_gpu_ctc_joint.py tests/test_gpu_turbo_parity.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/s3_pytest_decoder_pdl.log
timeout 600 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/s3_pytest_decoder.log
timeout 200 python tools/bench_decode.py 2>>gpurun_out/s3_decode.err | tee gpurun_out/s3_decode.json
DICOW_PDL=1 timeout 200 python tools/bench_decode.py 2>>gpurun_out/s3_decode.err | tee gpurun_out/s3_decode_pdl.json
DICOW_PDL=1 timeout 200 python tools/bench_decode.py --workload se_dicow 2>>gpurun_out/s3_decode.err | tee gpurun_out/s3_decode_se_pdl.json
timeout 200 python tools/bench_decode.py --workload se_dicow 2>>gpurun_out/s3_decode.err | tee gpurun_out/s3_decode_se.json
DICOW_PDL=1 timeout 200 python tools/profile_decode.py 2>&1 | grep -v -i warn | tail -62 > gpurun_out/s3_profile_decode_pdl.txt
