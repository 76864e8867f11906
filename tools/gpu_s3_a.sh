mkdir -p gpurun_out; rm -f gpurun_out/s3_decode_nt.json
timeout 600 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_beam.py tests/test_gpu_ctc_joint.py tests/test_gpu_kernels.py tests/test_gpu_turbo_parity.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/s3_pytest_decoder.log
for cfg in "" "--ln-kernel" "--ln-prologue" ""; do
echo "mode $cfg" | tee -a gpurun_out/s3_decode_nt.json
timeout 200 python tools/bench_decode.py $cfg 2>>gpurun_out/s3_decode.err | cut -c1-200 | tee -a gpurun_out/s3_decode_nt.json
done
timeout 200 python tools/profile_decode.py 2>&1 | grep -v -i warn | tail -62 > gpurun_out/s3_profile_decode_ln.txt
