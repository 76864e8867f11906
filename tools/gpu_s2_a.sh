#!/bin/bash
mkdir -p gpurun_out
python tools/profile_train.py --workload finetune > gpurun_out/s2_profile_train.txt 2>&1
head -50 gpurun_out/s2_profile_train.txt
