#!/usr/bin/env python
"""Diagnostic: the decode cross-attention kernel (one query per (batch, head) over 1500 keys, head-major K|V cache) at several batch
sizes: is a CTA's stream rate fixed (then 320 CTAs on 444 resident slots waste 28 %) or do the CTAs share the HBM rate?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ts_asr_whisper_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
H, T = 20, 1500
for B in (8, 16, 22, 32, 44, 64):
    g = torch.Generator(device=dev).manual_seed(B)
    # several cache copies so that consecutive launches do not hit L2 (B x 30.7 MB each; rotate over > 126 MB)
    n_copies = max(2, int(400e6 // (B * H * T * 128 * 2)) + 1)
    kvs = [(torch.randn(B, H, T, 128, device=dev, generator=g) * 0.5).bfloat16() for _ in range(n_copies)]
    q = (torch.randn(B, H * 64, device=dev, generator=g) * 0.3).bfloat16()
    out = torch.empty(B, H * 64, device=dev, dtype=torch.bfloat16)

    def run(i):
        kv = kvs[i % n_copies]
        ops.decode_attention(q, kv, kv[..., 64:], out, B=B, H=H, Tk=T, kv_row_stride=128, kv_batch_stride=H * T * 128,
                             kv_head_stride=T * 128)

    for i in range(5):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 40
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    gb = B * H * T * 128 * 2 / 1e9
    print(f"B={B:3d}: {B * H:5d} CTAs ({B * H / 148:5.2f} per SM)  {us:7.1f} us  {gb / us * 1e3:6.2f} TB/s")
