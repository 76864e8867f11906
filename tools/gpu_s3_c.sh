# decode recipes after the third session's linear-kernel work
mkdir -p gpurun_out; rm -f gpurun_out/s3_decode_recipes.json
timeout 300 python tools/bench_decode.py --workload se_dicow 2>>gpurun_out/s3_decode.err | tee -a gpurun_out/s3_decode_recipes.json | cut -c1-100
timeout 300 python tools/bench_decode.py --workload se_dicow --ctc-weight 0.2 2>>gpurun_out/s3_decode.err | tee -a gpurun_out/s3_decode_recipes.json | cut -c1-100
timeout 300 python tools/bench_decode.py --workload se_dicow --batch 12 --beams 5 --ctc-weight 0.2 2>>gpurun_out/s3_decode.err | tee -a gpurun_out/s3_decode_recipes.json | cut -c1-100
timeout 200 python tools/profile_decode.py 2>&1 | grep -v -i warn | tail -62 > gpurun_out/s3_profile_decode_final.txt
timeout 200 python tools/profile_decode.py --batch 12 --beams 5 --ctc-weight 0.2 --steps 16 2>&1 | grep -v -i warn | head -18 > gpurun_out/s3_profile_decode_beam.txt
