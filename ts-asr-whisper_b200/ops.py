"""torch-tensor wrappers around the C ABI: pointer/stride plumbing only, no arithmetic.

Every function enqueues CUDA work on torch's current stream of the tensors' device and returns immediately.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import lib as _lib
from .lib import (EPI_ACCUM_F32, EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GELU_BF16, EPI_DGELU_BF16, EPI_GELU_FDDT_POS_F32,
                  EPI_GELU_SAVE_BF16, EPI_RESIDUAL_F32, GEMM_A_T, GEMM_W_T, DicowError)

# number of kernels this module has launched (bench.py reports it as gpu_launches)
launch_count = 0

# Training-step scope for the prepared-weight caches (DiCoWEncoder.prepare / DiCoW.prepare_decoder): non-zero only INSIDE the
# forward / backward of one of training.py's autograd Functions.  A cache filled under epoch N is returned without recomputing
# its (data_ptr, version) key over every parameter while the epoch is N: the functions of one step ask for the prepared weights
# ~8 times, and the backward must see the forward's weights anyway.  Outside those scopes (0) every call checks the key.
prepare_epoch = 0
_epoch_counter = 0


def new_prepare_epoch() -> int:
    global _epoch_counter
    _epoch_counter += 1
    return _epoch_counter


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


class _Null:
    """no-op context manager (one shared instance)"""
    __slots__ = ()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NULL = _Null()


def _timed(kind: str, flops: float, dev: torch.device):
    return _NULL if timing_log is None else _Timed(kind, flops, dev)


def _guard(dev: torch.device):
    """device guard only when the tensor's device is not the current one (one process per GPU: practically never) -- a
    torch.cuda.device context costs two runtime calls per launch, and the fine-tune step issues ~1 500 launches from Python"""
    idx = dev.index
    if idx is None or torch.cuda.current_device() == idx:
        return _NULL
    return torch.cuda.device(dev)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class sm_budget:
    """``with ops.sm_budget(dev, n):`` -- the persistent kernels launched inside size their grids for ``n`` SMs
    (dicow_set_sm_budget), so that work on another stream keeps the rest of the device; restored on exit."""

    def __init__(self, dev: torch.device, sms: int):
        self.h, self.sms = _lib.handle(dev.index or 0), int(sms)

    def __enter__(self):
        self.effective = _lib.load_library().dicow_set_sm_budget(self.h, self.sms)
        return self

    def __exit__(self, *exc):
        _lib.load_library().dicow_set_sm_budget(self.h, 0)
        return False


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) if os.environ.get("DICOW_RAW_STREAM", "1") != "0" else None


def _stream(dev: torch.device) -> int:
    """cudaStream_t of torch's current stream on ``dev`` (the raw accessor skips building a torch.cuda.Stream object)"""
    if _raw_stream is not None and dev.index is not None:
        return _raw_stream(dev.index)
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
         pos: Optional[torch.Tensor] = None, flags: int = 0, splits: int = 0, N: Optional[int] = None,
         aux: Optional[torch.Tensor] = None, ldw: Optional[int] = None) -> torch.Tensor:
    """out[b, m, :] = epilogue(sum_k A[b, m, k] W[:, k]) -- see dicow_gemm_bf16 in include/dicow_b200.h.

    A: bf16, rows addressed as A + b*a_batch_stride + m*lda.  W: bf16 [N, K].  Defaults describe a plain
    contiguous 2-D GEMM (A [M, K], out [M, N]).
    """
    global launch_count
    dev = _require_cuda(A, W, out, bias, A2, resid, gate, stno, fddt_w, fddt_b, pos)
    assert A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
    if flags & GEMM_W_T:  # W holds Wt[k][n]
        Kw, Nw = W.shape
    else:
        Nw, Kw = W.shape
    N = Nw if N is None else N
    K = Kw if K is None else K
    if Mb is None:
        if flags & GEMM_A_T:  # A holds At[k][m]
            Mb = A.shape[-1]
        else:
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
    a.ldw = W.stride(0) if ldw is None else ldw
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
    a.flags = flags
    a.splits = splits
    a.aux_bf16 = _ptr(aux)
    h = _lib.handle(dev.index or 0)
    with _guard(dev), _timed("gemm", 2.0 * nb * Mb * N * K, dev):
        rc = _lib.load_library().dicow_gemm_bf16(h, C.byref(a), _stream(dev))
    _lib.check(rc, h, "dicow_gemm_bf16")
    launch_count += 1
    return out


def _call(name: str, dev: torch.device, args_struct, kind: str = "", flops: float = 0.0) -> None:
    global launch_count
    h = _lib.handle(dev.index or 0)
    with _guard(dev), _timed(kind or name, flops, dev):
        rc = getattr(_lib.load_library(), name)(h, C.byref(args_struct), _stream(dev))
    _lib.check(rc, h, name)
    launch_count += 1


def fddt_layernorm(x: torch.Tensor, *, T: int = 0, stno: Optional[torch.Tensor] = None,
                   fddt_w: Optional[torch.Tensor] = None, fddt_b: Optional[torch.Tensor] = None,
                   gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None, eps: float = 1e-5,
                   ln_out_bf16: Optional[torch.Tensor] = None, ln_out_f32: Optional[torch.Tensor] = None,
                   x_out_bf16: Optional[torch.Tensor] = None, delta1: Optional[torch.Tensor] = None,
                   delta2: Optional[torch.Tensor] = None, store_x: Optional[bool] = None, flags: int = 0,
                   x_out: Optional[torch.Tensor] = None) -> None:
    """x' = FDDT(x + delta1 + delta2) on the fp32 residual rows of ``x`` ([..., d], contiguous), written back when
    ``store_x`` (default: whenever x' differs from x), + LayerNorm outputs (dicow_fddt_layernorm).  ``stno`` is
    [B, 4, T] fp32 with ``B*T == rows``; the deltas are bf16 [rows, d] (pending out_proj / fc2 outputs)."""
    dev = _require_cuda(x, stno, fddt_w, fddt_b, gamma, beta, ln_out_bf16, ln_out_f32, x_out_bf16, delta1, delta2)
    for dl in (delta1, delta2):
        assert dl is None or (dl.dtype == torch.bfloat16 and dl.is_contiguous() and dl.numel() == x.numel())
    if store_x is None:
        store_x = stno is not None or delta1 is not None or delta2 is not None
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
    a.delta1_bf16, a.delta2_bf16 = _ptr(delta1), _ptr(delta2)
    a.store_x = 1 if store_x else 0
    a.flags = flags
    if x_out is not None:  # x' goes to x_out, x is left untouched
        assert store_x and x_out.dtype == torch.float32 and x_out.is_contiguous() and x_out.numel() == x.numel() and x_out.is_cuda
        a.x_out = _ptr(x_out)
    _call("dicow_fddt_layernorm", dev, a, "fddt_ln")


def fddt_full_combine(y: torch.Tensor, stno: torch.Tensor, x: torch.Tensor, *, T: int, pos: Optional[torch.Tensor] = None
                      ) -> torch.Tensor:
    """x[r] = sum_c stno[r // T, c, r % T] * y[r, c d:(c + 1) d] (+ pos[r % T]) -- the mask-weighted sum of the four class
    transforms of full-matrix FDDT (dicow_fddt_full_combine); y bf16 [rows, 4 d], x fp32 [rows, d]."""
    global launch_count
    dev = _require_cuda(y, stno, x, pos)
    rows, d = x.numel() // x.shape[-1], x.shape[-1]
    assert y.dtype == torch.bfloat16 and y.stride(-1) == 1 and x.dtype == torch.float32 and x.is_contiguous()
    assert stno.dtype == torch.float32 and stno.stride(2) == 1 and stno.stride(1) == stno.shape[2]
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_fddt_full_combine(h, _ptr(y), y.stride(-2), _ptr(stno), stno.stride(0), T, rows, d,
                                                         _ptr(pos), _ptr(x), _stream(dev))
    _lib.check(rc, h, "dicow_fddt_full_combine")
    launch_count += 1
    return x


def fddt_full_scatter(g: torch.Tensor, stno: torch.Tensor, dy: torch.Tensor, *, T: int) -> torch.Tensor:
    """dy[r, c d:(c + 1) d] = stno[r // T, c, r % T] * g[r] -- backward of fddt_full_combine (dicow_fddt_full_scatter);
    g fp32 [rows, d], dy bf16 [rows, 4 d]."""
    global launch_count
    dev = _require_cuda(g, stno, dy)
    rows, d = g.numel() // g.shape[-1], g.shape[-1]
    assert dy.dtype == torch.bfloat16 and dy.stride(-1) == 1 and g.dtype == torch.float32 and g.is_contiguous()
    assert stno.dtype == torch.float32 and stno.stride(2) == 1 and stno.stride(1) == stno.shape[2]
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_fddt_full_scatter(h, _ptr(g), _ptr(stno), stno.stride(0), T, rows, d, _ptr(dy),
                                                         dy.stride(-2), _stream(dev))
    _lib.check(rc, h, "dicow_fddt_full_scatter")
    launch_count += 1
    return dy


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, B: int, H: int, Tq: int,
              Tk: int, q_row_stride: int, q_batch_stride: int, kv_row_stride: int, kv_batch_stride: int,
              o_row_stride: int, o_batch_stride: int, causal: bool = False, variant: int = 0,
              lse: Optional[torch.Tensor] = None) -> torch.Tensor:
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
    a.lse = _ptr(lse)
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
    with _guard(dev):
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
    with _guard(dev):
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
    with _guard(dev):
        rc = _lib.load_library().dicow_cast_f32_bf16(h, _ptr(src), _ptr(out), src.numel(), _stream(dev))
    _lib.check(rc, h, "dicow_cast_f32_bf16")
    launch_count += 1
    return out


def stno_mask(activity: torch.Tensor, target: int, *, window_samples: int = 480000, frame_samples: int = 320,
              channels_first: bool = False) -> torch.Tensor:
    """STNO mask of one recording from per-speaker sample-level activity (dicow_stno_mask; reference
    src/data/local_datasets.py:162-196).  activity: bool / uint8 [n_speakers, n_samples] on the GPU; returns fp32
    [frames, 4] like the reference's get_stno_mask, or [4, frames] (the collated layout) with ``channels_first``."""
    global launch_count
    dev = _require_cuda(activity)
    act = activity.to(torch.uint8).contiguous()
    n_spk, n = act.shape
    frames = (n + (window_samples - n) % window_samples) // frame_samples
    out = torch.empty((4, frames) if channels_first else (frames, 4), dtype=torch.float32, device=dev)
    fs, cs = (1, frames) if channels_first else (4, 1)
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_stno_mask(h, _ptr(act), act.stride(0), n_spk, n, int(target), frame_samples, frames,
                                                 _ptr(out), fs, cs, _stream(dev))
    _lib.check(rc, h, "dicow_stno_mask")
    launch_count += 1
    return out


def augment_batch(stno: torch.Tensor, *, seg: Optional[torch.Tensor] = None, seg_soft: Optional[torch.Tensor] = None,
                  noise_rows: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                  feats: Optional[torch.Tensor] = None, factor: int = 2, spec: bool = False, warp: Optional[tuple] = None,
                  freq_masks: Optional[torch.Tensor] = None, time_masks: Optional[torch.Tensor] = None,
                  mask_channels: int = 128):
    """Apply a drawn augmentation plan to a padded batch (dicow_augment_batch; reference src/data/collators.py:184-210).
    stno fp32 [B, C, Ts] is changed in place by the segment / noise steps; with ``spec`` the SpecAug result is returned as
    new tensors (feats [B, M, factor * Ts], stno).  Index tables are int32, everything on the GPU."""
    dev = _require_cuda(stno, seg, seg_soft, noise_rows, noise, feats, freq_masks, time_masks)
    assert stno.dtype == torch.float32 and stno.is_contiguous() and stno.dim() == 3
    for t in (seg, noise_rows, freq_masks, time_masks):
        assert t is None or (t.dtype == torch.int32 and t.is_contiguous())
    for t in (seg_soft, noise, feats):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    a = _lib.AugmentArgs()
    a.struct_size = C.sizeof(_lib.AugmentArgs)
    a.stno = _ptr(stno)
    a.B, a.C, a.Ts = stno.shape
    a.seg, a.seg_soft, a.n_seg = _ptr(seg), _ptr(seg_soft), 0 if seg is None else seg.shape[0]
    a.noise_rows, a.noise, a.n_noise = _ptr(noise_rows), _ptr(noise), 0 if noise_rows is None else noise_rows.shape[0]
    assert noise is None or tuple(noise.shape) == (a.n_noise, a.C, a.Ts)
    a.spec = int(bool(spec))
    feats_out = stno_out = None
    if spec:
        assert feats is not None and feats.dim() == 3 and feats.shape[0] == a.B and feats.shape[2] == factor * a.Ts
        feats_out, stno_out = torch.empty_like(feats), torch.empty_like(stno)
        a.feats, a.feats_out, a.stno_out = _ptr(feats), _ptr(feats_out), _ptr(stno_out)
        a.M, a.Tf, a.factor = feats.shape[1], feats.shape[2], factor
        a.center, a.warped = (int(warp[0]), int(warp[1])) if warp is not None else (-1, -1)
        a.freq_masks, a.n_freq_masks = _ptr(freq_masks), 0 if freq_masks is None else freq_masks.shape[1]
        a.time_masks, a.n_time_masks = _ptr(time_masks), 0 if time_masks is None else time_masks.shape[1]
        a.mask_channels = mask_channels
    _call("dicow_augment_batch", dev, a, "augment")
    return (feats_out, stno_out) if spec else (feats, stno)


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


# ----------------------------------------------------------------------------------------------------------------
# decoder step / losses
# ----------------------------------------------------------------------------------------------------------------
def gemm_skinny(A: torch.Tensor, W: torch.Tensor, out: torch.Tensor, *, epilogue: int,
                bias: Optional[torch.Tensor] = None, M: Optional[int] = None, ldo: Optional[int] = None,
                resid: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None, pos_stride: int = 0
                ) -> torch.Tensor:
    """out[m, :] = epilogue(A[m, :] W^T + bias) for M <= 64 rows (dicow_gemm_skinny_bf16).  ``out`` may be a view
    into a larger buffer: its data_ptr(), ``ldo`` (row stride, elements) and ``pos`` (device int32 scalar: base
    advanced by pos * pos_stride elements) address the rows."""
    dev = _require_cuda(A, W, out, bias, resid, pos)
    a = _lib.GemmSkinnyArgs()
    a.struct_size = C.sizeof(_lib.GemmSkinnyArgs)
    a.A, a.lda = _ptr(A), A.stride(-2) if A.dim() > 1 else A.shape[-1]
    a.W, a.ldw = _ptr(W), W.stride(0)
    a.M = M if M is not None else A.numel() // A.shape[-1]
    a.N, a.K = W.shape
    a.bias = _ptr(bias)
    a.out = _ptr(out)
    a.ldo = ldo if ldo is not None else out.stride(-2)
    a.epilogue = epilogue
    a.resid = _ptr(resid)
    a.ldr = resid.stride(-2) if resid is not None else 0
    a.pos = _ptr(pos)
    a.pos_stride = pos_stride
    _call("dicow_gemm_skinny_bf16", dev, a, "gemm_skinny", 2.0 * a.M * a.N * a.K)
    return out


def decode_linear(W: torch.Tensor, out: torch.Tensor, *, epilogue: int, A: Optional[torch.Tensor] = None,
                  x: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
                  beta: Optional[torch.Tensor] = None, eps: float = 1e-5, bias: Optional[torch.Tensor] = None,
                  M: Optional[int] = None, ldo: Optional[int] = None, resid: Optional[torch.Tensor] = None,
                  out2: Optional[torch.Tensor] = None, n_split: int = 0, ldo2: int = 0,
                  pos: Optional[torch.Tensor] = None, pos_stride: int = 0) -> torch.Tensor:
    """[LayerNorm ->] Linear of one decode step in one kernel (dicow_decode_linear): the A operand is either
    LayerNorm(``x``) (fp32 [M, K], gamma / beta) or the bf16 ``A``; columns >= ``n_split`` go to ``out2`` (KV-cache
    append at device position ``pos``)."""
    src = x if x is not None else A
    dev = _require_cuda(src, W, out, gamma, beta, bias, resid, out2, pos)
    a = _lib.DecodeLinearArgs()
    a.struct_size = C.sizeof(_lib.DecodeLinearArgs)
    if x is not None:
        assert x.dtype == torch.float32 and gamma is not None and beta is not None
        a.x, a.ldx, a.gamma, a.beta, a.eps = _ptr(x), x.stride(-2) if x.dim() > 1 else x.shape[-1], _ptr(gamma), _ptr(beta), eps
    else:
        assert A is not None and A.dtype == torch.bfloat16
        a.A, a.lda = _ptr(A), A.stride(-2) if A.dim() > 1 else A.shape[-1]
    a.W, a.ldw = _ptr(W), W.stride(0)
    a.M = M if M is not None else src.numel() // src.shape[-1]
    a.N, a.K = W.shape
    a.bias = _ptr(bias)
    a.out = _ptr(out)
    a.ldo = ldo if ldo is not None else out.stride(-2)
    a.epilogue = epilogue
    a.resid = _ptr(resid)
    a.ldr = resid.stride(-2) if resid is not None else 0
    a.n_split, a.out2, a.ldo2 = n_split, _ptr(out2), ldo2
    a.pos, a.pos_stride = _ptr(pos), pos_stride
    _call("dicow_decode_linear", dev, a, "decode_linear", 2.0 * a.M * a.N * a.K)
    return out


def decode_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, B: int, H: int, Tk: int,
                     kv_row_stride: int, kv_batch_stride: int, pos: Optional[torch.Tensor] = None,
                     kv_head_stride: int = 0, kv_batch_div: int = 1, ancestry: Optional[torch.Tensor] = None) -> torch.Tensor:
    """one query row per (batch, head) against a K/V cache (dicow_decode_attention_bf16); q/out bf16 [B, H*64]."""
    dev = _require_cuda(q, k, v, out, pos)
    a = _lib.DecodeAttentionArgs()
    a.struct_size = C.sizeof(_lib.DecodeAttentionArgs)
    a.Q, a.q_batch_stride = _ptr(q), q.stride(0)
    a.K, a.V = _ptr(k), _ptr(v)
    a.kv_row_stride, a.kv_batch_stride = kv_row_stride, kv_batch_stride
    a.out, a.o_batch_stride = _ptr(out), out.stride(0)
    a.B, a.H, a.Tk = B, H, Tk
    a.pos = _ptr(pos)
    a.kv_head_stride = kv_head_stride
    a.kv_batch_div = kv_batch_div
    if ancestry is not None:
        assert ancestry.dtype == torch.int32 and ancestry.is_cuda and ancestry.stride(1) == 1
        a.ancestry, a.ancestry_stride = _ptr(ancestry), ancestry.stride(0)
    _call("dicow_decode_attention_bf16", dev, a, "decode_attention")
    return out


def kv_to_head_major(kv: torch.Tensor, out: torch.Tensor, *, B: int, T: int, H: int) -> torch.Tensor:
    """[B*T, (k | v) x H x 64] bf16 -> [B, H, T, 128] (dicow_kv_to_head_major): the cross-attention cache layout of the
    decode step (one contiguous K|V stream per (batch, head))."""
    global launch_count
    dev = _require_cuda(kv, out)
    assert kv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and kv.is_contiguous() and out.is_contiguous()
    assert kv.numel() == B * T * 2 * H * 64 == out.numel()
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_kv_to_head_major(h, _ptr(kv), _ptr(out), B, T, H, _stream(dev))
    _lib.check(rc, h, "dicow_kv_to_head_major")
    launch_count += 1
    return out


def embed_tokens(ids: torch.Tensor, tok: torch.Tensor, posw: torch.Tensor, x: torch.Tensor, *, S: int, past: int = 0,
                 pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x[b, s] = tok[ids[b, p + s]] + posw[p + s] with p = *pos or ``past`` (dicow_embed_tokens); fp32."""
    global launch_count
    dev = _require_cuda(ids, tok, posw, x, pos)
    assert ids.dtype == torch.int64 and tok.dtype == torch.float32 and posw.dtype == torch.float32
    assert tok.is_contiguous() and posw.is_contiguous() and x.is_contiguous() and x.dtype == torch.float32
    B = ids.shape[0]
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_embed_tokens(h, _ptr(ids), ids.stride(0), _ptr(tok), _ptr(posw), _ptr(x), B, S,
                                                    tok.shape[1], tok.shape[0], past, _ptr(pos), _stream(dev))
    _lib.check(rc, h, "dicow_embed_tokens")
    launch_count += 1
    return x


def decode_layer_table(layers: list, dev: torch.device) -> torch.Tensor:
    """device-resident dicow_decode_layer_args_t[L] for decode_layers: ``layers`` = one dict per decoder layer mapping the
    struct's field names to tensors (whose storage the caller keeps alive)"""
    arr = (_lib.DecodeLayerArgs * len(layers))()
    for i, entry in enumerate(layers):
        for name, _ in _lib.DecodeLayerArgs._fields_:
            t = entry[name]
            assert t.is_cuda and t.is_contiguous(), name
            setattr(arr[i], name, t.data_ptr())
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
    return raw.to(dev)


def decode_layers(table: torch.Tensor, *, B: int, d: int, H: int, ffn: int, L: int, T: int, S_max: int, vocab: int,
                  ids: torch.Tensor, tok: torch.Tensor, posw: torch.Tensor, pos: torch.Tensor, x: torch.Tensor,
                  q: torch.Tensor, ctx: torch.Tensor, hidden: torch.Tensor, barrier: torch.Tensor, workspace: torch.Tensor,
                  eps: float = 1e-5, flags: Optional[int] = None) -> None:
    """the decoder layers of one greedy token step in one persistent kernel (dicow_decode_layers)"""
    global launch_count
    dev = _require_cuda(table, ids, tok, posw, pos, x, q, ctx, hidden, barrier, workspace)
    assert workspace.dtype == torch.float32 and workspace.numel() >= B * H * 136
    assert ids.dtype == torch.int64 and tok.dtype == torch.float32 and posw.dtype == torch.float32 and x.dtype == torch.float32
    assert barrier.dtype == torch.int64 and pos.dtype == torch.int32
    a = _lib.DecodeLayersArgs()
    a.struct_size = C.sizeof(_lib.DecodeLayersArgs)
    a.B, a.d, a.H, a.ffn, a.L, a.T, a.S_max, a.vocab = B, d, H, ffn, L, T, S_max, vocab
    a.layers = _ptr(table)
    a.ids, a.ids_row_stride = _ptr(ids), ids.stride(0)
    a.embed_tokens, a.embed_positions, a.pos = _ptr(tok), _ptr(posw), _ptr(pos)
    a.x, a.q, a.ctx, a.hidden, a.barrier = _ptr(x), _ptr(q), _ptr(ctx), _ptr(hidden), _ptr(barrier)
    a.attn_workspace = _ptr(workspace)
    a.eps = eps
    a.flags = int(os.environ.get("DICOW_MEGA_FLAGS", "0")) if flags is None else flags
    _call("dicow_decode_layers", dev, a, "decode_layers")


def advance(pos: torch.Tensor, by: int = 1) -> None:
    global launch_count
    dev = _require_cuda(pos)
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_advance(h, _ptr(pos), by, _stream(dev))
    _lib.check(rc, h, "dicow_advance")
    launch_count += 1


def logits_rules_argmax(logits: torch.Tensor, ids: torch.Tensor, unfinished: torch.Tensor, *, begin_index: int, eos: int,
                        pad: int, no_timestamps: int, ts_begin: int, cur_len: int = 0,
                        pos: Optional[torch.Tensor] = None, max_initial_timestamp_index: Optional[int] = None,
                        suppress_bitmap: Optional[torch.Tensor] = None,
                        processed_scores: Optional[torch.Tensor] = None, timestamp_rules: bool = True,
                        no_select: bool = False) -> None:
    """suppress + Whisper timestamp rules + DiCoW EOS exception + argmax; appends the token to ``ids`` in place
    (dicow_logits_rules_argmax).  ``no_select``: only write ``processed_scores`` (joint CTC decoding selects later)."""
    dev = _require_cuda(logits, ids, unfinished, pos, suppress_bitmap, processed_scores)
    assert logits.dtype == torch.float32 and logits.stride(1) == 1 and ids.dtype == torch.int64
    assert unfinished.dtype == torch.int32
    a = _lib.LogitsRulesArgs()
    a.struct_size = C.sizeof(_lib.LogitsRulesArgs)
    a.logits, a.ld = _ptr(logits), logits.stride(0)
    a.B, a.V = logits.shape
    a.ids, a.ids_row_stride = _ptr(ids), ids.stride(0)
    a.pos, a.cur_len = _ptr(pos), cur_len
    a.begin_index, a.eos, a.pad, a.no_timestamps, a.ts_begin = begin_index, eos, pad, no_timestamps, ts_begin
    a.max_initial_timestamp_index = -1 if max_initial_timestamp_index is None else max_initial_timestamp_index
    a.timestamp_rules = 1 if timestamp_rules else 0
    a.suppress_bitmap = _ptr(suppress_bitmap)
    a.unfinished = _ptr(unfinished)
    a.processed_scores = _ptr(processed_scores)
    a.no_select = 1 if no_select else 0
    _call("dicow_logits_rules_argmax", dev, a, "logits_rules")


class CtcJointState:
    """Device buffers of joint CTC / attention decoding for one batch of hypotheses (dicow_ctc_joint_step):
    the window's CTC log-posteriors, the forward variables / prefix score of every hypothesis, candidate workspaces."""

    def __init__(self, ctc_logits: torch.Tensor, top_k: int = 500, upper_cased: Optional[dict] = None):
        assert ctc_logits.dim() == 3 and ctc_logits.dtype == torch.float32
        dev = _require_cuda(ctc_logits)
        B, T, V1 = ctc_logits.shape
        self.B, self.T, self.V1, self.K = B, T, V1, top_k
        self.logp = torch.empty(B, T, V1, dtype=torch.float32, device=dev)
        self.r_prev = torch.empty(B, T, 2, dtype=torch.float32, device=dev)
        self.score_prev = torch.empty(B, dtype=torch.float32, device=dev)
        self.states = torch.empty(B, T, 2, top_k, dtype=torch.float32, device=dev)
        self.ws_i32 = torch.zeros(4 * B + 4 + B * top_k, dtype=torch.int32, device=dev)
        self.ws_f32 = torch.zeros(B + 2 * B * top_k, dtype=torch.float32, device=dev)
        self._upper = None
        if upper_cased:  # decoding.py:183-186: an upper-cased token shares its lower-cased twin's posterior (column copy)
            self._upper = (torch.tensor(list(upper_cased.keys()), device=dev), torch.tensor(list(upper_cased.values()), device=dev))
        self.reset(ctc_logits)

    def reset(self, ctc_logits: torch.Tensor) -> None:
        """start a new window in the SAME buffers (captured CUDA graphs keep their pointers): posteriors of the window,
        initial forward variables r = (LOGZERO, running sum of the blank log-posteriors), score 0 (decoding.py:37-44,185)"""
        global launch_count
        assert tuple(ctc_logits.shape) == (self.B, self.T, self.V1) and ctc_logits.is_contiguous()
        dev = ctc_logits.device
        h = _lib.handle(dev.index or 0)
        with _guard(dev):
            rc = _lib.load_library().dicow_log_softmax_rows(h, _ptr(ctc_logits), _ptr(self.logp), self.B * self.T, self.V1,
                                                            _stream(dev))
        _lib.check(rc, h, "dicow_log_softmax_rows")
        launch_count += 1
        if self._upper is not None:
            self.logp[..., self._upper[1]] = self.logp[..., self._upper[0]]
        self.r_prev[:, :, 0] = -1e10
        self.r_prev[:, :, 1] = torch.cumsum(self.logp[:, :, self.V1 - 1], dim=1)
        self.score_prev.zero_()

    @property
    def candidates(self) -> torch.Tensor:
        return self.ws_i32[4 * self.B + 4:].view(self.B, self.K)

    @property
    def prefix_scores(self) -> torch.Tensor:
        return self.ws_f32[self.B + self.B * self.K:].view(self.B, self.K)


def ctc_joint_step(state: CtcJointState, processed_scores: torch.Tensor, ids: torch.Tensor, unfinished: torch.Tensor, *,
                   bos: int, eos: int, pad: int, first_timestamp: int, prefix_len: int, ctc_weight: float,
                   cur_len: int = 0, pos: Optional[torch.Tensor] = None, raw_logits: Optional[torch.Tensor] = None,
                   score_only: bool = False) -> None:
    """one token of joint CTC / attention greedy decoding for every hypothesis (dicow_ctc_joint_step): appends the token
    to ``ids`` in place, updates ``unfinished`` and the CTC state."""
    dev = _require_cuda(processed_scores, ids, unfinished, pos, state.logp, raw_logits)
    assert processed_scores.dtype == torch.float32 and processed_scores.is_contiguous() and ids.dtype == torch.int64
    assert unfinished.dtype == torch.int32 and processed_scores.shape[0] == state.B
    a = _lib.CtcJointArgs()
    a.struct_size = C.sizeof(_lib.CtcJointArgs)
    a.ids, a.ids_row_stride, a.pos, a.cur_len = _ptr(ids), ids.stride(0), _ptr(pos), cur_len
    a.B, a.V, a.T, a.V1, a.K = state.B, processed_scores.shape[1], state.T, state.V1, state.K
    a.bos, a.eos, a.pad, a.blank, a.first_timestamp, a.prefix_len = bos, eos, pad, max(state.V1 - 1, 0), first_timestamp, prefix_len
    a.ctc_weight = ctc_weight
    a.ctc_logp, a.processed_scores = _ptr(state.logp), _ptr(processed_scores)
    a.workspace_i32, a.workspace_f32, a.states = _ptr(state.ws_i32), _ptr(state.ws_f32), _ptr(state.states)
    a.r_prev, a.score_prev, a.unfinished = _ptr(state.r_prev), _ptr(state.score_prev), _ptr(unfinished)
    a.raw_logits = _ptr(raw_logits)
    a.score_only = 1 if score_only else 0
    _call("dicow_ctc_joint_step", dev, a, "ctc_joint")


class CandidateState:
    """workspaces of dicow_ctc_joint_step(score_only, ctc_weight = 0): attention-only beam search needs the log-softmax
    normaliser and the top-k candidates of every hypothesis, no CTC buffers"""

    def __init__(self, rows: int, top_k: int, device):
        self.B, self.K, self.T, self.V1 = rows, top_k, 0, 0
        self.logp = self.r_prev = self.score_prev = self.states = None
        self.ws_i32 = torch.zeros(4 * rows + 4 + rows * top_k, dtype=torch.int32, device=device)
        self.ws_f32 = torch.zeros(rows + 2 * rows * top_k, dtype=torch.float32, device=device)


def beam_step(*, U: int, NB: int, processed_scores: torch.Tensor, joint, ctc_weight: float, run_score: torch.Tensor,
              fin_score: torch.Tensor, fin_flag: torch.Tensor, unsat: torch.Tensor, ids: torch.Tensor, fin_ids: torch.Tensor,
              ids_tmp: torch.Tensor, ancestry: torch.Tensor, ancestry_tmp: torch.Tensor, pos: torch.Tensor, eos: int, pad: int,
              first_timestamp: int, max_length: int, prompt_len: int, length_penalty: float, early_stopping,
              scratch_i32: torch.Tensor, scratch_f32: torch.Tensor, flags: torch.Tensor,
              ctc_r_tmp: Optional[torch.Tensor] = None) -> None:
    """one step of beam search bookkeeping + re-linking on the device (dicow_beam_step); ``joint`` is the CtcJointState /
    CandidateState whose workspaces dicow_ctc_joint_step(score_only) just filled."""
    dev = _require_cuda(processed_scores, run_score, ids)
    a = _lib.BeamStepArgs()
    a.struct_size = C.sizeof(_lib.BeamStepArgs)
    a.U, a.NB, a.V, a.K = U, NB, processed_scores.shape[1], joint.K
    a.processed_scores = _ptr(processed_scores)
    a.joint_workspace_i32, a.joint_workspace_f32 = _ptr(joint.ws_i32), _ptr(joint.ws_f32)
    a.ctc_weight = ctc_weight
    if ctc_weight > 0:
        a.ctc_states, a.ctc_r_prev, a.ctc_score_prev = _ptr(joint.states), _ptr(joint.r_prev), _ptr(joint.score_prev)
        a.ctc_r_tmp, a.T = _ptr(ctc_r_tmp), joint.T
    a.run_score, a.fin_score, a.fin_flag, a.unsat = _ptr(run_score), _ptr(fin_score), _ptr(fin_flag), _ptr(unsat)
    a.ids, a.fin_ids, a.ids_tmp, a.ids_row_stride = _ptr(ids), _ptr(fin_ids), _ptr(ids_tmp), ids.stride(0)
    a.ancestry, a.ancestry_tmp, a.ancestry_stride = _ptr(ancestry), _ptr(ancestry_tmp), ancestry.stride(0)
    a.pos, a.eos, a.pad, a.first_timestamp = _ptr(pos), eos, pad, first_timestamp
    a.max_length, a.prompt_len, a.length_penalty = max_length, prompt_len, length_penalty
    a.early_stopping = 2 if early_stopping == "never" else (1 if early_stopping is True else 0)
    a.scratch_i32, a.scratch_f32, a.flags = _ptr(scratch_i32), _ptr(scratch_f32), _ptr(flags)
    _call("dicow_beam_step", dev, a, "beam_step")


def suppress_bitmap(token_ids, vocab: int, device) -> torch.Tensor:
    """int32 bitmap (bit v of word v // 32 set = token v suppressed) for logits_rules_argmax."""
    words = [0] * ((vocab + 31) // 32)
    for t in token_ids or ():
        if 0 <= int(t) < vocab:
            words[int(t) >> 5] |= 1 << (int(t) & 31)
    words = [w - (1 << 32) if w >= (1 << 31) else w for w in words]
    return torch.tensor(words, dtype=torch.int32, device=device)


def softlabel_ce(logits: torch.Tensor, labels: torch.Tensor, upp_labels: Optional[torch.Tensor] = None, *,
                 ts_begin: int = 0, smoothing: Optional[torch.Tensor] = None, soft_mode: bool = True) -> torch.Tensor:
    """Decoder loss (dicow_softlabel_ce): logits fp32 [rows, V]; labels int64 [rows].  Returns a 0-dim fp32 tensor."""
    dev = _require_cuda(logits, labels, upp_labels, smoothing)
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    labels = labels.reshape(-1).contiguous()
    upp = upp_labels.reshape(-1).contiguous() if upp_labels is not None else None
    rows, V = logits.shape
    assert labels.numel() == rows and labels.dtype == torch.int64
    ws = torch.empty(2 * rows, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    a = _lib.SoftlabelCeArgs()
    a.struct_size = C.sizeof(_lib.SoftlabelCeArgs)
    a.logits, a.ld, a.rows, a.V = _ptr(logits), logits.stride(0), rows, V
    a.labels, a.upp_labels = _ptr(labels), _ptr(upp)
    a.ts_begin = ts_begin
    a.n_ts = smoothing.shape[0] if smoothing is not None else 0
    a.smoothing = _ptr(smoothing)
    a.soft_mode = 1 if soft_mode else 0
    a.workspace, a.loss = _ptr(ws), _ptr(loss)
    _call("dicow_softlabel_ce", dev, a, "softlabel_ce")
    return loss


def _ctc_rows(logits: torch.Tensor) -> torch.Tensor:
    """[B, T, V1] fp32 logits as the CTC kernels address them: unit column stride, rows ``stride(1)`` elements apart and batches
    T rows apart (a row-padded buffer, as the CTC head writes it, passes through; anything else is made contiguous)"""
    if logits.dim() == 3 and logits.stride(2) == 1 and logits.stride(1) >= logits.shape[2] and \
            (logits.shape[0] == 1 or logits.stride(0) == logits.shape[1] * logits.stride(1)):
        return logits
    return logits.contiguous()


def ctc_loss(logits: torch.Tensor, labels: torch.Tensor, reduction: str = "mean") -> torch.Tensor:
    """CTC loss with blank = last class, all frames valid, zero_infinity (dicow_ctc_loss).  logits fp32 [B, T, V+1]
    contiguous, labels int64 [B, Lmax] (negative = padding)."""
    dev = _require_cuda(logits, labels)
    assert logits.dtype == torch.float32 and labels.dtype == torch.int64
    logits = _ctc_rows(logits)
    if reduction not in ("mean", "sum"):
        raise NotImplementedError(f"ctc_loss_reduction={reduction}")
    labels = labels.contiguous()
    B, T, V1 = logits.shape
    ws = torch.empty(B * T + 2 * B, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    a = _lib.CtcLossArgs()
    a.struct_size = C.sizeof(_lib.CtcLossArgs)
    a.logits, a.B, a.T, a.V1 = _ptr(logits), B, T, V1
    a.labels, a.Lmax = _ptr(labels), labels.shape[1]
    a.reduction_mean = 1 if reduction == "mean" else 0
    a.workspace, a.loss = _ptr(ws), _ptr(loss)
    a.ld = logits.stride(1)
    _call("dicow_ctc_loss", dev, a, "ctc_loss")
    return loss


# ----------------------------------------------------------------------------------------------------------------
# backward (training)
# ----------------------------------------------------------------------------------------------------------------
def attention_bwd(q, k, v, o, do, lse, dq, dk, dv, *, B: int, H: int, Tq: int, Tk: int, q_row_stride: int,
                  q_batch_stride: int, kv_row_stride: int, kv_batch_stride: int, o_row_stride: int, o_batch_stride: int,
                  dq_row_stride: int, dq_batch_stride: int, dkv_row_stride: int, dkv_batch_stride: int,
                  causal: bool = False) -> None:
    """dQ / dK / dV of softmax(Q K^T) V (dicow_attention_bwd_bf16); all tensors bf16 views addressed by explicit
    strides (elements), ``lse`` fp32 [B, H, Tq] as saved by attention(..., lse=...); ``do`` shares ``o``'s strides."""
    dev = _require_cuda(q, k, v, o, do, lse, dq, dk, dv)
    # per-query D = sum(dO * O), plus (non-causal encoder shapes) the fp32 dQ accumulator of the single-pass kernel
    n_ws = B * H * Tq * 65 + 4 if (not causal and Tq >= 256 and Tk >= 128) else B * H * Tq
    ws = torch.empty(n_ws, dtype=torch.float32, device=dev)
    a = _lib.AttentionBwdArgs()
    a.struct_size = C.sizeof(_lib.AttentionBwdArgs)
    a.Q, a.K, a.V, a.O, a.dO, a.lse = _ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(do), _ptr(lse)
    a.dQ, a.dK, a.dV = _ptr(dq), _ptr(dk), _ptr(dv)
    a.B, a.H, a.Tq, a.Tk = B, H, Tq, Tk
    a.q_row_stride, a.q_batch_stride = q_row_stride, q_batch_stride
    a.kv_row_stride, a.kv_batch_stride = kv_row_stride, kv_batch_stride
    a.o_row_stride, a.o_batch_stride = o_row_stride, o_batch_stride
    a.dq_row_stride, a.dq_batch_stride = dq_row_stride, dq_batch_stride
    a.dkv_row_stride, a.dkv_batch_stride = dkv_row_stride, dkv_batch_stride
    a.causal = 1 if causal else 0
    a.workspace = _ptr(ws)
    a.workspace_floats = n_ws
    fl = 10.0 * B * H * Tq * Tk * 64 * (0.5 if causal else 1.0)  # 5 GEMMs; S and dP are recomputed in the 2nd pass
    _call("dicow_attention_bwd_bf16", dev, a, "attention_bwd", fl)


def layernorm_fddt_bwd(x: torch.Tensor, g_out: torch.Tensor, *, dy: Optional[torch.Tensor] = None,
                       g_in: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None, eps: float = 1e-5,
                       delta1: Optional[torch.Tensor] = None, delta2: Optional[torch.Tensor] = None, T: int = 0,
                       stno: Optional[torch.Tensor] = None, fddt_w: Optional[torch.Tensor] = None,
                       fddt_b: Optional[torch.Tensor] = None, g_out_bf16: Optional[torch.Tensor] = None,
                       dgamma: Optional[torch.Tensor] = None, dbeta: Optional[torch.Tensor] = None,
                       dfddt_w: Optional[torch.Tensor] = None, dfddt_b: Optional[torch.Tensor] = None,
                       g_colsum: Optional[torch.Tensor] = None) -> None:
    """backward of fddt_layernorm (dicow_layernorm_fddt_bwd); parameter gradients are accumulated (+=).  ``g_colsum`` [d] fp32
    += the column sums of ``g_out_bf16`` (the bias gradient of the Linear whose output was a pending delta)."""
    dev = _require_cuda(x, g_out, dy, g_in, gamma, delta1, delta2, stno, fddt_w, fddt_b, g_out_bf16, dgamma, dbeta,
                        dfddt_w, dfddt_b, g_colsum)
    a = _lib.LnBwdArgs()
    a.struct_size = C.sizeof(_lib.LnBwdArgs)
    a.x, a.delta1_bf16, a.delta2_bf16 = _ptr(x), _ptr(delta1), _ptr(delta2)
    a.d = x.shape[-1]
    a.rows = x.numel() // x.shape[-1]
    a.T = T if T else a.rows
    a.stno = _ptr(stno)
    a.stno_batch_stride = stno.stride(0) if stno is not None else 0
    a.fddt_w, a.fddt_b = _ptr(fddt_w), _ptr(fddt_b)
    a.gamma, a.eps = _ptr(gamma), eps
    a.dy_bf16, a.g_in, a.g_out, a.g_out_bf16 = _ptr(dy), _ptr(g_in), _ptr(g_out), _ptr(g_out_bf16)
    a.dgamma, a.dbeta, a.dfddt_w, a.dfddt_b = _ptr(dgamma), _ptr(dbeta), _ptr(dfddt_w), _ptr(dfddt_b)
    if g_colsum is not None:
        assert g_colsum.dtype == torch.float32 and g_colsum.numel() == x.shape[-1] and g_out_bf16 is not None
        a.g_colsum = _ptr(g_colsum)
    _call("dicow_layernorm_fddt_bwd", dev, a, "ln_bwd")


def colsum(x: torch.Tensor, out: torch.Tensor, alpha: float = 1.0) -> None:
    """out[n] += alpha * sum_rows x[row, n] (dicow_colsum); x bf16 / fp32 [rows, N] with unit column stride."""
    global launch_count
    dev = _require_cuda(x, out)
    assert x.dim() == 2 and x.stride(1) == 1 and out.dtype == torch.float32 and x.dtype in (torch.bfloat16, torch.float32)
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_colsum(h, _ptr(x), 1 if x.dtype == torch.bfloat16 else 0, x.stride(0), x.shape[0],
                                              out.numel(), _ptr(out), alpha, _stream(dev))
    _lib.check(rc, h, "dicow_colsum")
    launch_count += 1


def conv1d_col2im(dcol: torch.Tensor, dx: torch.Tensor, *, B: int, T: int, T_out: int, C_in: int, stride: int,
                  dx_batch_stride: int, dx_row_stride: int) -> None:
    global launch_count
    dev = _require_cuda(dcol, dx)
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_conv1d_col2im(h, _ptr(dcol), _ptr(dx), B, T, T_out, C_in, stride, dx_batch_stride,
                                                     dx_row_stride, _stream(dev))
    _lib.check(rc, h, "dicow_conv1d_col2im")
    launch_count += 1


def ctc_loss_with_lse(logits: torch.Tensor, labels: torch.Tensor, reduction: str = "mean"):
    """CTC loss value plus the per-row log-sum-exp workspace the backward consumes (dicow_ctc_loss)."""
    dev = _require_cuda(logits, labels)
    assert logits.dtype == torch.float32 and labels.dtype == torch.int64
    logits = _ctc_rows(logits)
    if reduction not in ("mean", "sum"):
        raise NotImplementedError(f"ctc_loss_reduction={reduction}")
    labels = labels.contiguous()
    B, T, V1 = logits.shape
    ws_f = torch.empty(B * T + 2 * B, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    a = _lib.CtcLossArgs()
    a.struct_size = C.sizeof(_lib.CtcLossArgs)
    a.logits, a.B, a.T, a.V1 = _ptr(logits), B, T, V1
    a.labels, a.Lmax = _ptr(labels), labels.shape[1]
    a.reduction_mean = 1 if reduction == "mean" else 0
    a.workspace, a.loss = _ptr(ws_f), _ptr(loss)
    a.ld = logits.stride(1)
    _call("dicow_ctc_loss", dev, a, "ctc_loss")
    return loss, ws_f


def ctc_loss_bwd(logits: torch.Tensor, labels: torch.Tensor, lse_ws: torch.Tensor, reduction: str = "mean",
                 loss_scale: float = 1.0, scale_dev: Optional[torch.Tensor] = None, out_f32: bool = False) -> torch.Tensor:
    """loss_scale * (*scale_dev) * dL/dlogits (dicow_ctc_loss_bwd): bf16 [B, T, ceil8(V1)] with zeroed padding columns
    (the wgrad GEMM's operand), or fp32 [B, T, V1] when ``out_f32`` (what autograd hands to a visible logits tensor)."""
    dev = _require_cuda(logits, labels, lse_ws, scale_dev)
    labels = labels.contiguous()
    logits = _ctc_rows(logits)
    B, T, V1 = logits.shape
    Lmax = labels.shape[1]
    S = 2 * Lmax + 1
    ldd = V1 if out_f32 else -(-V1 // 8) * 8
    ws_b = torch.empty(2 * B * T * S + B, dtype=torch.float32, device=dev)
    dlogits = torch.empty(B, T, ldd, dtype=torch.float32 if out_f32 else torch.bfloat16, device=dev)
    b = _lib.CtcBwdArgs()
    b.struct_size = C.sizeof(_lib.CtcBwdArgs)
    b.logits, b.lse, b.B, b.T, b.V1 = _ptr(logits), _ptr(lse_ws), B, T, V1
    b.labels, b.Lmax, b.reduction_mean, b.loss_scale = _ptr(labels), Lmax, 1 if reduction == "mean" else 0, loss_scale
    b.workspace, b.dlogits_bf16, b.ldd = _ptr(ws_b), _ptr(dlogits), ldd
    b.scale_dev, b.out_f32 = _ptr(scale_dev), 1 if out_f32 else 0
    b.ld = logits.stride(1)
    _call("dicow_ctc_loss_bwd", dev, b, "ctc_bwd")
    return dlogits


def ctc_loss_fwd_bwd(logits: torch.Tensor, labels: torch.Tensor, reduction: str = "mean", loss_scale: float = 1.0):
    """CTC loss value and loss_scale * dL/dlogits (bf16, rows padded to a multiple of 8 columns for the wgrad GEMM)."""
    loss, ws_f = ctc_loss_with_lse(logits, labels, reduction)
    return loss, ctc_loss_bwd(logits, labels, ws_f, reduction, loss_scale)


def softlabel_ce_bwd(logits: torch.Tensor, labels: torch.Tensor, upp_labels: Optional[torch.Tensor] = None, *,
                     ts_begin: int = 0, smoothing: Optional[torch.Tensor] = None, soft_mode: bool = True,
                     scale: float = 1.0, scale_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """scale * (*scale_dev) * d(sum of per-token losses)/dlogits as bf16 [rows, ceil8(V)] (dicow_softlabel_ce_bwd)."""
    dev = _require_cuda(logits, labels, upp_labels, smoothing, scale_dev)
    labels = labels.reshape(-1).contiguous()
    upp = upp_labels.reshape(-1).contiguous() if upp_labels is not None else None
    rows, V = logits.shape
    ldd = -(-V // 8) * 8
    out = torch.empty(rows, ldd, dtype=torch.bfloat16, device=dev)
    a = _lib.SoftlabelCeBwdArgs()
    a.struct_size = C.sizeof(_lib.SoftlabelCeBwdArgs)
    a.logits, a.ld, a.rows, a.V = _ptr(logits), logits.stride(0), rows, V
    a.labels, a.upp_labels = _ptr(labels), _ptr(upp)
    a.ts_begin = ts_begin
    a.n_ts = smoothing.shape[0] if smoothing is not None else 0
    a.smoothing = _ptr(smoothing)
    a.soft_mode = 1 if soft_mode else 0
    a.scale = scale
    a.scale_dev = _ptr(scale_dev)
    a.dlogits_bf16, a.ldd = _ptr(out), ldd
    _call("dicow_softlabel_ce_bwd", dev, a, "ce_bwd")
    return out


def dgelu_mul(g: torch.Tensor, pre: torch.Tensor) -> torch.Tensor:
    """bf16 [rows, cols] = g * gelu'(pre) (dicow_dgelu_mul); g fp32 / bf16 and pre bf16, 2-D with unit column stride."""
    global launch_count
    dev = _require_cuda(g, pre)
    assert g.dim() == 2 and pre.shape == g.shape and g.stride(1) == 1 and pre.stride(1) == 1
    assert pre.dtype == torch.bfloat16 and g.dtype in (torch.bfloat16, torch.float32)
    out = torch.empty(g.shape, dtype=torch.bfloat16, device=dev)
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_dgelu_mul(h, _ptr(g), 1 if g.dtype == torch.bfloat16 else 0, g.stride(0), _ptr(pre),
                                                 pre.stride(0), _ptr(out), out.stride(0), g.shape[0], g.shape[1], _stream(dev))
    _lib.check(rc, h, "dicow_dgelu_mul")
    launch_count += 1
    return out


def gate_bwd(g: torch.Tensor, upd: torch.Tensor, gate: torch.Tensor, dgate: Optional[torch.Tensor]) -> torch.Tensor:
    """SE-DiCoW gate backward (dicow_gate_bwd): returns bf16 tanh(gate) * g and accumulates the scalar gate gradient
    (1 - tanh^2) * sum(g * upd) into ``dgate`` (fp32 [1] or None).  g fp32 [rows, cols], upd bf16 [rows, cols]."""
    global launch_count
    dev = _require_cuda(g, upd, gate, dgate)
    assert g.dim() == 2 and upd.shape == g.shape and g.stride(1) == 1 and upd.stride(1) == 1
    assert g.dtype == torch.float32 and upd.dtype == torch.bfloat16 and gate.dtype == torch.float32
    out = torch.empty(g.shape, dtype=torch.bfloat16, device=dev)
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_gate_bwd(h, _ptr(g), g.stride(0), _ptr(upd), upd.stride(0), _ptr(gate), _ptr(out),
                                                out.stride(0), g.shape[0], g.shape[1], _ptr(dgate), _stream(dev))
    _lib.check(rc, h, "dicow_gate_bwd")
    launch_count += 1
    return out


def embedding_bwd(g: torch.Tensor, ids: torch.Tensor, *, S: int, d_tok: Optional[torch.Tensor],
                  d_pos: Optional[torch.Tensor], past: int = 0) -> None:
    """d_tok[ids[r]] += g[r], d_pos[past + r % S] += g[r] (dicow_embedding_bwd); g fp32 [B * S, d] contiguous."""
    global launch_count
    dev = _require_cuda(g, ids, d_tok, d_pos)
    assert g.dtype == torch.float32 and g.is_contiguous() and ids.dtype == torch.int64 and ids.is_contiguous()
    rows, d = g.shape
    vocab = d_tok.shape[0] if d_tok is not None else 0
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_embedding_bwd(h, _ptr(g), _ptr(ids), rows, S, d, past, vocab, _ptr(d_tok), _ptr(d_pos),
                                                     _stream(dev))
    _lib.check(rc, h, "dicow_embedding_bwd")
    launch_count += 1


def cast_bf16_padded(src: torch.Tensor, cols_out: int) -> torch.Tensor:
    """fp32 [rows, cols] (row stride arbitrary) -> bf16 [rows, cols_out] with zero padding (dicow_cast_f32_bf16_2d)."""
    global launch_count
    dev = _require_cuda(src)
    assert src.dim() == 2 and src.dtype == torch.float32 and src.stride(1) == 1 and cols_out >= src.shape[1]
    out = torch.empty(src.shape[0], cols_out, dtype=torch.bfloat16, device=dev)
    h = _lib.handle(dev.index or 0)
    with _guard(dev):
        rc = _lib.load_library().dicow_cast_f32_bf16_2d(h, _ptr(src), src.stride(0), _ptr(out), cols_out, src.shape[0],
                                                        src.shape[1], cols_out, _stream(dev))
    _lib.check(rc, h, "dicow_cast_f32_bf16_2d")
    launch_count += 1
    return out
