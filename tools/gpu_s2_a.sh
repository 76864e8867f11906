#!/bin/bash
mkdir -p gpurun_out
python tools/profile_train.py --workload finetune > gpurun_out/s2_profile_train2.txt 2>&1
head -42 gpurun_out/s2_profile_train2.txt | cut -c1-150
