"""Energy per launch of the hot kernels (NVML total-energy counter), measured by looping ONE kernel for ~2 s.

The encoder step runs at the board power cap (sw_power_cap, ~1.0 kW, SM clock 1.45-1.6 GHz of 1.965): the step time is then
(energy per step) / (cap), so the quantity to minimise per kernel is JOULES per launch, not isolated milliseconds (a kernel
timed alone boosts to ~1.9 GHz).  Prints, per kernel: ms / launch when looped alone, average W, J / launch, SM MHz, and
"ms at 1 kW" = J / 1000 W -- what the kernel costs inside a power-capped step.

    python tools/energy_probe.py [--seconds 2.0] [--only attn,gemm,ln]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml  # noqa: E402

from ts_asr_whisper_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--only", default="")
args = ap.parse_args()
dev = torch.device("cuda:0")
pynvml.nvmlInit()
nv = pynvml.nvmlDeviceGetHandleByIndex(0)
B, T, H, d, ffn = 32, 1500, 20, 1280, 5120
rows = B * T
g = torch.Generator(device=dev).manual_seed(0)


def measure(name, fn, flops=0.0, nbytes=0.0):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    # calibrate launches for ~args.seconds
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    n = max(50, int(args.seconds * 1e3 / (e0.elapsed_time(e1) / 20)))
    time.sleep(0.3)
    torch.cuda.synchronize()
    j0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(nv)
    t0 = time.perf_counter()
    e0.record()
    clocks = []
    for i in range(n):
        fn()
        if i % max(1, n // 8) == 0:
            clocks.append(pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM))
    e1.record()
    torch.cuda.synchronize()
    j1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(nv)
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / n
    joule = (j1 - j0) * 1e-3 / n
    extra = ""
    if flops:
        extra += f"  {flops / ms / 1e9:7.1f} TFLOP/s alone, {flops / joule / 1e12:5.2f} TFLOP/J"
    if nbytes:
        extra += f"  {nbytes / ms / 1e6:7.1f} GB/s alone"
    print(f"{name:46s} {ms * 1e3:8.1f} us alone  {joule / wall * n:6.0f} W  {joule * 1e3:8.2f} mJ/launch  "
          f"{sorted(clocks)[len(clocks) // 2]:5d} MHz  -> {joule:8.5f} ms at 1 kW x1000={joule * 1e3:7.1f} us{extra}", flush=True)


def want(k):
    return not args.only or k in args.only.split(",")


if want("attn"):
    qkv = (torch.randn(B, T, 3 * d, device=dev, generator=g) * 0.5).bfloat16()
    out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16)
    fl = 4.0 * B * H * T * T * 64
    for variant in (0, 1, 2, 4):
        measure(f"attention variant {variant}", lambda v=variant: ops.attention(
            qkv, qkv[:, :, d:], qkv[:, :, 2 * d:], out, B=B, H=H, Tq=T, Tk=T, q_row_stride=3 * d, q_batch_stride=T * 3 * d,
            kv_row_stride=3 * d, kv_batch_stride=T * 3 * d, o_row_stride=d, o_batch_stride=T * d, variant=v), fl)
    q = qkv[:, :, :d].view(B, T, H, 64).transpose(1, 2)
    k = qkv[:, :, d:2 * d].view(B, T, H, 64).transpose(1, 2)
    v = qkv[:, :, 2 * d:].view(B, T, H, 64).transpose(1, 2)
    measure("torch SDPA (library, context)", lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0), fl)
    del qkv, out

if want("gemm"):
    A = (torch.randn(rows, d, device=dev, generator=g) * 0.5).bfloat16()
    A4 = (torch.randn(rows, ffn, device=dev, generator=g) * 0.5).bfloat16()
    for name, (a, N, K, epi) in {"qkv  [48000x1280]x[3840]  bias->bf16": (A, 3 * d, d, ops.EPI_BIAS_BF16),
                                 "out  [48000x1280]x[1280]  bias->bf16": (A, d, d, ops.EPI_BIAS_BF16),
                                 "fc1  [48000x1280]x[5120]  bias+GELU->bf16": (A, ffn, d, ops.EPI_BIAS_GELU_BF16),
                                 "fc1  [48000x1280]x[5120]  bias->bf16 (no GELU)": (A, ffn, d, ops.EPI_BIAS_BF16),
                                 "fc2  [48000x5120]x[1280]  bias->bf16": (A4, d, ffn, ops.EPI_BIAS_BF16)}.items():
        W = (torch.randn(N, K, device=dev, generator=g) * 0.03).bfloat16()
        bias = torch.randn(N, device=dev, generator=g)
        o = torch.empty(rows, N, device=dev, dtype=torch.bfloat16)
        measure("gemm " + name, lambda: ops.gemm(a, W, o, epilogue=epi, bias=bias), 2.0 * rows * N * K)
        measure("     same, single-CTA kernel (flags=1)", lambda: ops.gemm(a, W, o, epilogue=epi, bias=bias, flags=1), 2.0 * rows * N * K)
        Af, Wf = a, W
        measure("     torch.matmul (cuBLAS, context)", lambda: torch.matmul(Af, Wf.t()), 2.0 * rows * N * K)
        del W, o
    del A, A4

if want("ln"):
    x = torch.randn(rows, d, device=dev, generator=g)
    d1 = torch.randn(rows, d, device=dev, generator=g).bfloat16()
    d2 = torch.randn(rows, d, device=dev, generator=g).bfloat16()
    ln = torch.empty(rows, d, device=dev, dtype=torch.bfloat16)
    stno = torch.softmax(torch.randn(B, 4, T, device=dev, generator=g), 1)
    fw, fb = torch.rand(4, d, device=dev) + 0.5, torch.randn(4, d, device=dev) * 0.1
    gam, bet = torch.rand(d, device=dev) + 0.5, torch.randn(d, device=dev) * 0.1
    measure("fddt+ln1 (x += d1 + d2, FDDT, LN -> bf16): 14 B/el", lambda: ops.fddt_layernorm(
        x, T=T, stno=stno, fddt_w=fw, fddt_b=fb, gamma=gam, beta=bet, ln_out_bf16=ln, delta1=d1, delta2=d2), 0, 14.0 * rows * d)
    measure("ln2 (LN(x + d1) -> bf16, x not stored): 8 B/el", lambda: ops.fddt_layernorm(
        x, gamma=gam, beta=bet, ln_out_bf16=ln, delta1=d1, store_x=False), 0, 8.0 * rows * d)
