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
    with torch.cuda.device(dev):
        rc = _lib.load_library().dicow_gemm_bf16(h, C.byref(a), _stream(dev))
    _lib.check(rc, h, "dicow_gemm_bf16")
    launch_count += 1
    return out
