"""Timing probe for dicow_attention_bf16 (both P-staging variants) at the benchmark shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ts_asr_whisper_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def run(B, H, T, variant, iters=10):
    d = H * 64
    qkv = (torch.randn(B, T, 3 * d, device=dev) * 0.5).bfloat16()
    out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16)

    def call():
        ops.attention(qkv, qkv[:, :, d:], qkv[:, :, 2 * d:], out, B=B, H=H, Tq=T, Tk=T, q_row_stride=3 * d,
                      q_batch_stride=T * 3 * d, kv_row_stride=3 * d, kv_batch_stride=T * 3 * d, o_row_stride=d,
                      o_batch_stride=T * d, variant=variant)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * B * H * T * T * 64
    print(f"attention B={B} H={H} T={T} variant={variant}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    if variant != 0:
        return
    q = qkv[:, :, :d].view(B, T, H, 64).transpose(1, 2)
    k = qkv[:, :, d:2 * d].view(B, T, H, 64).transpose(1, 2)
    v = qkv[:, :, 2 * d:].view(B, T, H, 64).transpose(1, 2)
    for _ in range(3):
        torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"   torch SDPA: {ms2:.3f} ms  {fl / ms2 / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    for variant in (0, 1, 2, 3, 4):
        try:
            run(32, 20, 1500, variant)
        except Exception as ex:  # noqa: BLE001
            print(f"variant {variant} failed: {ex}")
