#!/bin/bash
python tools/profile_train.py --workload ctc_pretrain 2>&1 | grep -E "256, 3|cast_2d|total GPU|device span"
python tools/profile_train.py --workload finetune 2>&1 | grep -E "256, 3|cast_2d|total GPU|device span"
