mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_beam.py tests/test_gpu_ctc_joint.py -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/s3_pytest_decoder.log
timeout 300 python tools/bench_decode.py --workload se_dicow --batch 12 --beams 5 --ctc-weight 0.2 2>>gpurun_out/s3_decode.err | tee gpurun_out/s3_decode_beam_nt.json | cut -c1-100
timeout 200 python tools/profile_decode.py --batch 12 --beams 5 --ctc-weight 0.2 --steps 16 2>&1 | grep -v -i warn | head -18 > gpurun_out/s3_profile_decode_beam_nt.txt
