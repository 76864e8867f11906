#!/bin/bash
python tools/profile_train.py --workload finetune 2>&1 | head -8 | cut -c1-260
