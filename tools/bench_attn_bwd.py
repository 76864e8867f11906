#!/usr/bin/env python
"""Diagnostic: attention backward at the fine-tune step's shape (B = 8, H = 20, T = 1500, head_dim 64), looped alone.
DICOW_ATTN_BWD_FUSED=0/1 selects the two-pass / single-pass kernels.  Not a bench value."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ts_asr_whisper_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, H, T = int(os.environ.get("B", "8")), 20, 1500
d = H * 64
g = torch.Generator(device=dev).manual_seed(0)
qkv = (torch.randn(B, T, 3 * d, device=dev, generator=g) * 0.5).bfloat16()
do = (torch.randn(B, T, d, device=dev, generator=g) * 0.5).bfloat16()
out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device=dev)
kw = dict(B=B, H=H, Tq=T, Tk=T, q_row_stride=3 * d, q_batch_stride=T * 3 * d, kv_row_stride=3 * d, kv_batch_stride=T * 3 * d,
          o_row_stride=d, o_batch_stride=T * d)
ops.attention(qkv, qkv[:, :, d:], qkv[:, :, 2 * d:], out, lse=lse, **kw)
dqkv = torch.empty_like(qkv)


def run():
    ops.attention_bwd(qkv, qkv[:, :, d:], qkv[:, :, 2 * d:], out, do, lse, dqkv, dqkv[:, :, d:], dqkv[:, :, 2 * d:],
                      dq_row_stride=3 * d, dq_batch_stride=T * 3 * d, dkv_row_stride=3 * d, dkv_batch_stride=T * 3 * d, **kw)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 10.0 * B * H * T * T * 64
print(f"attention_bwd fused={os.environ.get('DICOW_ATTN_BWD_FUSED', '1')} B={B}: {ms * 1e3:.1f} us, {fl / ms / 1e9:.0f} TFLOP/s (5 GEMMs)")
