#!/bin/bash
mkdir -p gpurun_out
for k in "test_ctc_pretrain_step or megakernel" "updates_and_second or unfreeze or megakernel" "se_dicow_finetune or megakernel"; do
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_turbo_parity.py -q -x -k "$k" > gpurun_out/s2_c.log 2>&1; echo "[$k] rc=$?"; head -1 gpurun_out/s2_c.log | cut -c1-100; tail -1 gpurun_out/s2_c.log | cut -c1-100
done
