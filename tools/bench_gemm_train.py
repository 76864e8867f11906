#!/usr/bin/env python
"""Diagnostic: the GEMM shapes of one encoder layer of the fine-tune step (B = 8: 12 000 rows), forward / dgrad / wgrad, on the
single-CTA kernel (flags 1) and on CTA pairs (flags 2); each looped alone with CUDA events.  Not a bench value."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ts_asr_whisper_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = int(os.environ.get("ROWS", "12000"))


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.1).bfloat16()
    for (N, K, name) in [(3840, 1280, "qkv"), (1280, 1280, "out"), (5120, 1280, "fc1"), (1280, 5120, "fc2")]:
        X, W, dY = rnd(M, K), rnd(N, K), rnd(M, N)
        b = torch.randn(N, device=dev, generator=g)
        Y, pre = torch.empty(M, N, dtype=torch.bfloat16, device=dev), torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        dX = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
        dW = torch.zeros(N, K, device=dev)
        gf = 2.0 * M * N * K / 1e9
        row = [f"{name:4s} M={M} N={N} K={K}"]
        for form in (1, 2):
            t_f = timeit(lambda: ops.gemm(X, W, Y, epilogue=ops.EPI_BIAS_BF16, bias=b, flags=form))
            t_s = timeit(lambda: ops.gemm(X, W, Y, epilogue=ops.EPI_GELU_SAVE_BF16, bias=b, aux=pre, flags=form))
            t_d = timeit(lambda: ops.gemm(dY, W, dX, epilogue=ops.EPI_BIAS_BF16, flags=ops.GEMM_W_T | form))
            t_g = timeit(lambda: ops.gemm(dY, W, dX, epilogue=ops.EPI_DGELU_BF16, flags=ops.GEMM_W_T | form, aux=X)) \
                if name == "fc1" or True else 0
            t_w = timeit(lambda: ops.gemm(dY, X, dW, epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_A_T | ops.GEMM_W_T | form))
            row.append(f"form{form}: fwd {t_f*1e3:6.1f} us {gf/t_f:6.0f} TF | gelu_save {t_s*1e3:6.1f} {gf/t_s:6.0f} | dgrad {t_d*1e3:6.1f} "
                       f"{gf/t_d:6.0f} | dgelu {t_g*1e3:6.1f} {gf/t_g:6.0f} | wgrad {t_w*1e3:6.1f} {gf/t_w:6.0f}")
        print("\n   ".join(row), flush=True)
    t = timeit(lambda: torch.matmul(X, W.t()))
    print(f"torch.matmul last shape fwd: {t*1e3:.1f} us")


if __name__ == "__main__":
    main()
