# third session of round 2, last call: the whole GPU suite on the final library, then compute-sanitizer (memcheck + racecheck) on
# the kernels the session changed (decode linear, LayerNorm rows, rules + argmax)
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/s3_pytest_gpu_final.log
timeout 40 compute-sanitizer --tool racecheck --kernel-name "regex=decode_linear_kernel|logits_rules_argmax" --error-exitcode 9 \
    python -m pytest tests/test_gpu_decoder.py -m gpu -q -x -k "test_decode_linear or rules" > gpurun_out/s3_sanitizer_race.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/s3_sanitizer_race.log; tail -3 gpurun_out/s3_sanitizer_race.log
