#!/usr/bin/env python
"""Training-step benchmarks of the hot path (BASELINE.json configs[2] and configs[4]) as a command line; the same functions
run inside bench.py (its "secondary" object carries them into the driver's record).  One JSON line per run.

    python tools/bench_train.py --workload finetune      [--batch 8]  [--steps K] [--warmup W]
    python tools/bench_train.py --workload ctc_pretrain  [--batch 16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_train.py --workload finetune --gpus N
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="finetune", choices=["finetune", "ctc_pretrain"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--labels", type=int, default=64)
    ap.add_argument("--no-overlap", action="store_true", help="exchange gradients after the backward instead of during it")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from bench import ClockSampler
    from tools import workloads
    from ts_asr_whisper_b200 import parallel
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    parallel.init_process_group("nccl", dev)
    sampler = None
    if rank == 0:
        sampler = ClockSampler(local_rank)
        sampler.start()
    out = workloads.train_step(args.workload, dev, rank, world, steps=args.steps, warmup=args.warmup, batch=args.batch,
                               label_len=args.labels, overlap=not args.no_overlap, sampler=sampler)
    if rank == 0:
        sampler.stop()
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
