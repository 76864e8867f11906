"""Time the GPU collator augmentations at the fine-tune recipe's batch shape (8 x 128 x 3000 mel frames, 8 x 4 x 1500 STNO):
device time of dicow_augment_batch (CUDA events) and the whole collator call (host padding into the pinned staging buffer,
H2D, plan draw, kernels).   python tools/bench_augment.py  ->  one JSON line"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ts_asr_whisper_b200 import ops  # noqa: E402
from ts_asr_whisper_b200.collators import DataCollator  # noqa: E402


class Tok:
    upper_cased_tokens = {}

    def __call__(self, texts, **_):
        class E(dict):
            attention_mask = torch.ones(len(texts), 3, dtype=torch.long)
        return E(input_ids=torch.arange(3).repeat(len(texts), 1) + 5)


def main():
    B, M, Tf = 8, 128, 3000
    rng = np.random.default_rng(1)
    ins = []
    for _ in range(B):
        raw = rng.random((Tf // 2, 4)).astype(np.float32) + np.float32(1e-3)
        ins.append({"is_long_form": False, "transcript": "x", "input_features": torch.from_numpy(rng.standard_normal((M, Tf)).astype(np.float32)),
                    "attention_mask": torch.ones(Tf, dtype=torch.long), "stno_mask": torch.from_numpy(raw / raw.sum(1, keepdims=True))})
    col = DataCollator(feature_extractor=None, tokenizer=Tok(), bos_token_id=0, max_length=16, stno_gaussian_noise_var=0.002,
                       stno_gaussian_noise_prob=1.0, stno_segment_augment_prob=1.0, spec_aug_prob=1.0, device="cuda")
    torch.manual_seed(0)
    for _ in range(3):
        col(ins)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        batch = col(ins)
    torch.cuda.synchronize()
    call_ms = (time.perf_counter() - t0) / n * 1e3
    feats, stno = batch["input_features"].clone(), batch["stno_mask"].clone()
    plan = col.draw_plan(B, 4, Tf // 2, M, Tf).to("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    times = []
    for _ in range(10):
        flush.zero_()
        s = stno.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        col.apply_plan(feats, s, plan)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    dev_us = float(np.median(times)) * 1e3
    bytes_moved = 2 * (feats.numel() + stno.numel()) * 4
    print(json.dumps({"workload": f"collator augmentations, batch {B} x {M} x {Tf} + STNO {B} x 4 x {Tf // 2}, all three on",
                      "device_us": round(dev_us, 1), "algorithmic_GBps": round(bytes_moved / dev_us / 1e3, 1),
                      "collator_call_ms": round(call_ms, 2), "segments": 0 if plan.seg is None else int(plan.seg.shape[0]),
                      "l2": "flushed between launches", "launches_per_call": 3, "ops_launch_count": ops.launch_count}))


if __name__ == "__main__":
    main()
