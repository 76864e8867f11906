"""torch-tensor wrappers around the C ABI: pointer/stride plumbing only, no arithmetic.

Every function enqueues CUDA work on torch's current stream of the tensors' device and returns immediately.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as _lib
from .lib import (EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GELU_BF16, EPI_GELU_FDDT_POS_F32, EPI_RESIDUAL_F32,
                  DicowError)

# number of kernels this module has launched (bench.py reports it as gpu_launches)
launch_count = 0

# Optional per-launch timing hook (bench.py's roofline leg): when set to a list, gemm()/attention() append
# (kind, flops, start_event, end_event) with CUDA events recorded on the launching stream around the launch.
timing_log = None


class _Timed:
    __slots__ = ("kind", "flops", "dev", "e0")

    def __init__(self, kind: str, flops: float, dev: torch.device):
        self.kind, self.flops, self.dev = kind, flops, dev

    def __enter__(self):
        if timing_log is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.dev))
        return self

    def __exit__(self, *exc):
        if timing_log is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(torch.cuda.current_stream(self.dev))
            timing_log.append((self.kind, self.flops, self.e0, e1))
        return False


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _require_cuda(*ts: torch.Tensor) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise DicowError("dicow ops run on CUDA tensors only (no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise DicowError("tensors on different devices")
    return dev


def gemm(A: torch.Tensor, W: torch.Tensor, out: torch.Tensor, *, epilogue: int, bias: Optional[torch.Tensor] = None,
         nb: int = 1, Mb: Optional[int] = None, K: Optional[int] = None, lda: Optional[int] = None,
         a_batch_stride: int = 0, ldo: Optional[int] = None, out_batch_stride: int = 0,
         A2: Optional[torch.Tensor] = None, lda2: int = 0, a2_batch_stride: int = 0, K1: int = 0,
         resid: Optional[torch.Tensor] = None, ldr: int = 0, resid_batch_stride: int = 0,
         gate: Optional[torch.Tensor] = None, stno: Optional[torch.Tensor] = None, stno_batch_stride: int = 0,
         fddt_w: Optional[torch.Tensor] = None, fddt_b: Optional[torch.Tensor] = None,
         pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[b, m, :] = epilogue(sum_k A[b, m, k] W[:, k]) -- see dicow_gemm_bf16 in include/dicow_b200.h.

    A: bf16, rows addressed as A + b*a_batch_stride + m*lda.  W: bf16 [N, K].  Defaults describe a plain
    contiguous 2-D GEMM (A [M, K], out [M, N]).
    """
    global launch_count
    dev = _require_cuda(A, W, out, bias, A2, resid, gate, stno, fddt_w, fddt_b, pos)
    assert A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
    N, Kw = W.shape
    K = Kw if K is None else K
    if Mb is None:
        Mb = A.numel() // A.shape[-1] if nb == 1 else A.shape[-2]
    a = _lib.GemmArgs()
    a.struct_size = C.sizeof(_lib.GemmArgs)
    a.A = _ptr(A)
    a.lda = A.stride(-2) if lda is None else lda
    a.a_batch_stride = a_batch_stride
    a.A2 = _ptr(A2)
    a.lda2 = lda2
    a.a2_batch_stride = a2_batch_stride
    a.K1 = K1
    a.W = _ptr(W)
    a.ldw = W.stride(0)
    a.nb, a.Mb, a.N, a.K = nb, Mb, N, K
    a.bias = _ptr(bias)
    a.out = _ptr(out)
    a.ldo = out.stride(-2) if ldo is None else ldo
    a.out_batch_stride = out_batch_stride
    a.epilogue = epilogue
    a.resid = _ptr(resid)
    a.ldr = ldr if ldr else (resid.stride(-2) if resid is not None else 0)
    a.resid_batch_stride = resid_batch_stride
    a.gate = _ptr(gate)
    a.stno = _ptr(stno)
    a.stno_batch_stride = stno_batch_stride
    a.fddt_w = _ptr(fddt_w)
    a.fddt_b = _ptr(fddt_b)
    a.pos = _ptr(pos)
    h = _lib.handle(dev.index or 0)
    with torch.cuda.device(dev), _Timed("gemm", 2.0 * nb * Mb * N * K, dev):
        rc = _lib.load_library().dicow_gemm_bf16(h, C.byref(a), _stream(dev))
    _lib.check(rc, h, "dicow_gemm_bf16")
    launch_count += 1
    return out


def _call(name: str, dev: torch.device, args_struct, kind: str = "", flops: float = 0.0) -> None:
    global launch_count
    h = _lib.handle(dev.index or 0)
    with torch.cuda.device(dev), _Timed(kind or name, flops, dev):
        rc = getattr(_lib.load_library(), name)(h, C.byref(args_struct), _stream(dev))
    _lib.check(rc, h, name)
    launch_count += 1


def fddt_layernorm(x: torch.Tensor, *, T: int = 0, stno: Optional[torch.Tensor] = None,
                   fddt_w: Optional[torch.Tensor] = None, fddt_b: Optional[torch.Tensor] = None,
                   gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None, eps: float = 1e-5,
                   ln_out_bf16: Optional[torch.Tensor] = None, ln_out_f32: Optional[torch.Tensor] = None,
                   x_out_bf16: Optional[torch.Tensor] = None) -> None:
    """In-place FDDT on the fp32 residual rows of ``x`` ([..., d], contiguous) + LayerNorm outputs
    (dicow_fddt_layernorm).  ``stno`` is [B, 4, T] fp32 with ``B*T == rows``."""
    dev = _require_cuda(x, stno, fddt_w, fddt_b, gamma, beta, ln_out_bf16, ln_out_f32, x_out_bf16)
    assert x.dtype == torch.float32 and x.is_contiguous()
    a = _lib.FddtLnArgs()
    a.struct_size = C.sizeof(_lib.FddtLnArgs)
    a.x = _ptr(x)
    a.d = x.shape[-1]
    a.rows = x.numel() // x.shape[-1]
    a.T = T if T else a.rows
    a.stno = _ptr(stno)
    a.stno_batch_stride = stno.stride(0) if stno is not None else 0
    if stno is not None:
        assert stno.dtype == torch.float32 and stno.stride(2) == 1 and stno.stride(1) == stno.shape[2]
    a.fddt_w = _ptr(fddt_w)
    a.fddt_b = _ptr(fddt_b)
    a.gamma = _ptr(gamma)
    a.beta = _ptr(beta)
    a.eps = eps
    a.ln_out_bf16 = _ptr(ln_out_bf16)
    a.ln_out_f32 = _ptr(ln_out_f32)
    a.x_out_bf16 = _ptr(x_out_bf16)
    _call("dicow_fddt_layernorm", dev, a, "fddt_ln")


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, B: int, H: int, Tq: int,
              Tk: int, q_row_stride: int, q_batch_stride: int, kv_row_stride: int, kv_batch_stride: int,
              o_row_stride: int, o_batch_stride: int, causal: bool = False, variant: int = 0) -> torch.Tensor:
    """softmax(Q K^T) V per (batch, head), head_dim 64 (dicow_attention_bf16).  q/k/v/out may be views into fused
    buffers: only their data_ptr() and the explicit strides (elements) are used."""
    dev = _require_cuda(q, k, v, out)
    a = _lib.AttentionArgs()
    a.struct_size = C.sizeof(_lib.AttentionArgs)
    a.Q, a.K, a.V, a.out = _ptr(q), _ptr(k), _ptr(v), _ptr(out)
    a.B, a.H, a.Tq, a.Tk = B, H, Tq, Tk
    a.q_row_stride, a.q_batch_stride = q_row_stride, q_batch_stride
    a.kv_row_stride, a.kv_batch_stride = kv_row_stride, kv_batch_stride
    a.o_row_stride, a.o_batch_stride = o_row_stride, o_batch_stride
    a.causal = 1 if causal else 0
    a.variant = variant
    # algorithmic FLOPs: QK^T + PV = 4 * Tq * Tk * 64 per (batch, head); causal counts the visible half
    fl = 4.0 * B * H * Tq * Tk * 64 * (0.5 if causal else 1.0)
    _call("dicow_attention_bf16", dev, a, "attention", fl)
    return out


def features_to_channels_last(feats: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """fp32 [B, C, F] -> zero-padded channels-last bf16 [B, F + 2, C]."""
    global launch_count
    dev = _require_cuda(feats, out)
    B, Cc, F = feats.shape
    assert feats.dtype == torch.float32 and feats.is_contiguous() and out.dtype == torch.bfloat16
    assert out.shape == (B, F + 2, Cc) and out.is_contiguous()
    h = _lib.handle(dev.index or 0)
    with torch.cuda.device(dev):
        rc = _lib.load_library().dicow_features_to_channels_last(h, _ptr(feats), _ptr(out), B, Cc, F, _stream(dev))
    _lib.check(rc, h, "dicow_features_to_channels_last")
    launch_count += 1
    return out


def zero_pad_rows(buf: torch.Tensor) -> None:
    """zero rows 0 and T+1 of a channels-last bf16 [B, T + 2, C] buffer."""
    global launch_count
    dev = _require_cuda(buf)
    B, Tp, Cc = buf.shape
    h = _lib.handle(dev.index or 0)
    with torch.cuda.device(dev):
        rc = _lib.load_library().dicow_zero_pad_rows(h, _ptr(buf), B, Tp - 2, Cc, _stream(dev))
    _lib.check(rc, h, "dicow_zero_pad_rows")
    launch_count += 1


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 copy through the library's cast kernel (weight preparation)."""
    global launch_count
    dev = _require_cuda(src)
    src = src.contiguous()
    out = torch.empty(src.shape, dtype=torch.bfloat16, device=dev)
    h = _lib.handle(dev.index or 0)
    with torch.cuda.device(dev):
        rc = _lib.load_library().dicow_cast_f32_bf16(h, _ptr(src), _ptr(out), src.numel(), _stream(dev))
    _lib.check(rc, h, "dicow_cast_f32_bf16")
    launch_count += 1
    return out


def logmel(audio: torch.Tensor, mel_filters: torch.Tensor, lengths: Optional[torch.Tensor] = None,
           return_attention_mask: bool = False):
    """Whisper log-mel of a batch of zero-padded recordings (dicow_logmel): audio fp32 [B, n_pad] (n_pad % 160 == 0),
    mel_filters fp32 [201, M], lengths int64 [B].  Returns input_features [B, M, n_pad // 160] fp32 (and the int32
    attention mask [B, n_pad // 160])."""
    dev = _require_cuda(audio, mel_filters, lengths)
    assert audio.dtype == torch.float32 and audio.dim() == 2 and audio.stride(1) == 1
    assert mel_filters.dtype == torch.float32 and mel_filters.is_contiguous() and mel_filters.shape[0] == 201
    B, n_pad = audio.shape
    M = mel_filters.shape[1]
    frames = n_pad // 160
    out = torch.empty(B, M, frames, dtype=torch.float32, device=dev)
    mask = torch.empty(B, frames, dtype=torch.int32, device=dev) if return_attention_mask else None
    if return_attention_mask and lengths is None:
        lengths = torch.full((B,), n_pad, dtype=torch.int64, device=dev)
    if lengths is not None:
        assert lengths.dtype == torch.int64 and lengths.is_contiguous()
    ws = torch.empty(B, dtype=torch.int32, device=dev)
    a = _lib.LogmelArgs()
    a.struct_size = C.sizeof(_lib.LogmelArgs)
    a.audio, a.audio_batch_stride, a.B, a.n_pad = _ptr(audio), audio.stride(0), B, n_pad
    a.lengths, a.mel_filters, a.n_mels = _ptr(lengths), _ptr(mel_filters), M
    a.out, a.attention_mask, a.workspace = _ptr(out), _ptr(mask), _ptr(ws)
    _call("dicow_logmel", dev, a, "logmel")
    return (out, mask) if return_attention_mask else out
