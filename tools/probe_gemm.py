"""GPU bring-up probe for dicow_gemm_bf16: each case compared with torch fp32 matmul of the same bf16 inputs.

usage: python tools/probe_gemm.py [case_index ...]   (no args = all, sequentially in this process)
Each case prints one line; on mismatch it prints where the errors are (tile/row/col structure) to help
debug descriptor / swizzle mistakes offline.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ts_asr_whisper_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def describe_err(out, ref, name):
    err = (out.float() - ref.float()).abs()
    tol = 2e-2 * ref.float().abs().max().item() + 1e-3
    bad = err > tol
    print(f"  [{name}] max_abs_err={err.max().item():.4e} ref_absmax={ref.float().abs().max().item():.3e} "
          f"bad={bad.sum().item()}/{bad.numel()} nan={torch.isnan(out.float()).sum().item()}")
    if bad.any():
        idx = bad.nonzero()
        rows = idx[:, -2].unique()
        cols = idx[:, -1].unique()
        print(f"    bad rows: n={rows.numel()} first={rows[:16].tolist()} last={rows[-4:].tolist()}")
        print(f"    bad cols: n={cols.numel()} first={cols[:16].tolist()} last={cols[-4:].tolist()}")
        r, c = idx[0, -2].item(), idx[0, -1].item()
        o2 = out.reshape(-1, out.shape[-2], out.shape[-1])
        r2 = ref.reshape(-1, ref.shape[-2], ref.shape[-1])
        b = idx[0, 0].item() if idx.shape[1] == 3 else 0
        print(f"    first bad at b={b} r={r} c={c}: got {o2[b, r, c:c+8].float().tolist()} want {r2[b, r, c:c+8].float().tolist()}")
    return not bad.any().item()


def gelu(x):
    return torch.nn.functional.gelu(x)


def case_plain(M, N, K, epi=ops.EPI_BIAS_BF16, bias=True, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    A = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=dev, generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device=dev, generator=g) if bias else None
    ref = A.float() @ W.float().t()
    if bias:
        ref = ref + b
    if epi == ops.EPI_BIAS_BF16:
        out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.gemm(A, W, out, epilogue=epi, bias=b)
    elif epi == ops.EPI_BIAS_GELU_BF16:
        ref = gelu(ref)
        out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.gemm(A, W, out, epilogue=epi, bias=b)
    elif epi == ops.EPI_BIAS_F32:
        out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
        ops.gemm(A, W, out, epilogue=epi, bias=b)
    elif epi == ops.EPI_RESIDUAL_F32:
        res = torch.randn(M, N, device=dev, generator=g)
        gate = torch.tensor([0.7], device=dev)
        ref = res + torch.tanh(gate) * ref
        out = res.clone()
        ops.gemm(A, W, out, epilogue=epi, bias=b, resid=out, gate=gate)
    torch.cuda.synchronize()
    return describe_err(out, ref, f"plain M={M} N={N} K={K} epi={epi}")


def case_conv(B, T, Cin, Cout, stride, seed=1):
    """Conv1d(k=3, pad=1) as a GEMM over a zero-padded channels-last buffer with overlapping rows."""
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, Cin, T, device=dev, generator=g) * 0.5
    w = torch.randn(Cout, Cin, 3, device=dev, generator=g) * 0.05
    b = torch.randn(Cout, device=dev, generator=g)
    xb = x.bfloat16().float()
    wb = w.bfloat16().float()
    ref = gelu(torch.nn.functional.conv1d(xb, wb, b, stride=stride, padding=1)).transpose(1, 2).contiguous()
    Tout = ref.shape[1]
    xp = torch.zeros(B, T + 2, Cin, device=dev, dtype=torch.bfloat16)
    xp[:, 1:T + 1] = x.transpose(1, 2).bfloat16()
    W2 = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous().bfloat16()  # [n, k*Cin + c]
    out = torch.full((B, Tout, Cout), float("nan"), device=dev, dtype=torch.bfloat16)
    ops.gemm(xp, W2, out, epilogue=ops.EPI_BIAS_GELU_BF16, bias=b, nb=B, Mb=Tout, K=3 * Cin, lda=stride * Cin,
             a_batch_stride=(T + 2) * Cin, ldo=Cout, out_batch_stride=Tout * Cout)
    torch.cuda.synchronize()
    return describe_err(out, ref, f"conv B={B} T={T} Cin={Cin} Cout={Cout} s={stride}")


def case_split(M, N, K1, K2, seed=2):
    g = torch.Generator(device=dev).manual_seed(seed)
    A1 = (torch.randn(M, K1, device=dev, generator=g) * 0.5).bfloat16()
    A2 = (torch.randn(M, K2, device=dev, generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K1 + K2, device=dev, generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device=dev, generator=g)
    ref = gelu(torch.cat([A1, A2], -1).float() @ W.float().t() + b)
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    ops.gemm(A1, W, out, epilogue=ops.EPI_BIAS_GELU_BF16, bias=b, K=K1 + K2, A2=A2, lda2=K2, K1=K1)
    torch.cuda.synchronize()
    return describe_err(out, ref, f"splitK M={M} N={N} K1={K1} K2={K2}")


def case_fddt(B, T, Cin, Cout, seed=3):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, Cin, 2 * T, device=dev, generator=g) * 0.5
    w = torch.randn(Cout, Cin, 3, device=dev, generator=g) * 0.05
    b = torch.randn(Cout, device=dev, generator=g)
    stno = torch.softmax(3 * torch.randn(B, 4, T, device=dev, generator=g), dim=1)
    fw = torch.rand(4, Cout, device=dev, generator=g) + 0.5
    fb = torch.randn(4, Cout, device=dev, generator=g) * 0.1
    pos = torch.randn(T, Cout, device=dev, generator=g) * 0.1
    y = gelu(torch.nn.functional.conv1d(x.bfloat16().float(), w.bfloat16().float(), b, stride=2, padding=1)).transpose(1, 2)
    ref = sum((y * fw[c] + fb[c]) * stno[:, c, :, None] for c in range(4)) + pos
    xp = torch.zeros(B, 2 * T + 2, Cin, device=dev, dtype=torch.bfloat16)
    xp[:, 1:2 * T + 1] = x.transpose(1, 2).bfloat16()
    W2 = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous().bfloat16()
    out = torch.full((B, T, Cout), float("nan"), device=dev, dtype=torch.float32)
    ops.gemm(xp, W2, out, epilogue=ops.EPI_GELU_FDDT_POS_F32, bias=b, nb=B, Mb=T, K=3 * Cin, lda=2 * Cin,
             a_batch_stride=(2 * T + 2) * Cin, ldo=Cout, out_batch_stride=T * Cout, stno=stno,
             stno_batch_stride=4 * T, fddt_w=fw, fddt_b=fb, pos=pos)
    torch.cuda.synchronize()
    return describe_err(out, ref, f"conv2+fddt+pos B={B} T={T} Cin={Cin} Cout={Cout}")


def timing(M, N, K, epi=ops.EPI_BIAS_BF16, iters=20, flags=0):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    if epi == ops.EPI_RESIDUAL_F32:
        out = torch.zeros(M, N, device=dev, dtype=torch.float32)
        kw = dict(resid=out)
    else:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw = {}
    for _ in range(3):
        ops.gemm(A, W, out, epilogue=epi, bias=b, flags=flags, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(A, W, out, epilogue=epi, bias=b, flags=flags, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS reference for context
    for _ in range(3):
        torch.matmul(A, W.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(A, W.t())
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print(f"  [timing M={M} N={N} K={K} epi={epi} flags={flags}] {ms:.3f} ms  {tf:.1f} TFLOP/s   (cuBLAS {ms2:.3f} ms "
          f"{2.0 * M * N * K / ms2 / 1e9:.1f} TFLOP/s)")
    return True


CASES = [
    lambda: case_plain(128, 256, 64, bias=False),
    lambda: case_plain(128, 256, 256, bias=False),
    lambda: case_plain(128, 128, 128),
    lambda: case_plain(256, 512, 256),
    lambda: case_plain(1500, 1280, 1280),
    lambda: case_plain(3000, 5120, 1280, epi=ops.EPI_BIAS_GELU_BF16),
    lambda: case_plain(3000, 1280, 5120, epi=ops.EPI_RESIDUAL_F32),
    lambda: case_plain(777, 1003, 320, epi=ops.EPI_BIAS_F32),
    lambda: case_plain(100, 384, 384, epi=ops.EPI_BIAS_BF16),
    lambda: case_plain(20000, 3840, 1280),
    lambda: case_conv(2, 3000, 128, 1280, 1),
    lambda: case_conv(2, 3000, 1280, 1280, 2),
    lambda: case_conv(3, 200, 64, 128, 2),
    lambda: case_split(1500, 5120, 1280, 1280),
    lambda: case_fddt(2, 1500, 256, 1280),
    lambda: timing(48000, 3840, 1280),
    lambda: timing(48000, 5120, 1280, epi=ops.EPI_BIAS_GELU_BF16),
    lambda: timing(48000, 1280, 5120, epi=ops.EPI_RESIDUAL_F32),
    lambda: timing(48000, 1280, 1280),
    lambda: timing(48000, 3840, 1280, flags=2),
    lambda: timing(48000, 5120, 1280, epi=ops.EPI_BIAS_GELU_BF16, flags=2),
    lambda: timing(48000, 1280, 5120, epi=ops.EPI_RESIDUAL_F32, flags=2),
    lambda: timing(48000, 1280, 1280, flags=2),
    lambda: timing(48000, 3840, 1280, flags=1),
    lambda: timing(48000, 5120, 1280, epi=ops.EPI_BIAS_GELU_BF16, flags=1),
]

if __name__ == "__main__":
    sel = [int(a) for a in sys.argv[1:]] or list(range(len(CASES)))
    ok_all = True
    for i in sel:
        t0 = time.time()
        try:
            ok = CASES[i]()
        except Exception as ex:  # noqa: BLE001
            ok = False
            print(f"  case {i} raised: {type(ex).__name__}: {ex}")
        print(f"case {i}: {'PASS' if ok else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
        ok_all &= bool(ok)
    sys.exit(0 if ok_all else 1)
