"""Training step (forward with saved activations + hand-scheduled backward) of the DiCoW hot path on the C-ABI kernels.

The reference trains through HF Trainer: ``loss = model(**batch).loss; loss.backward()`` under bf16 autocast with fp32
master parameters (src/utils/trainers.py:116-139, configs/base.yaml:49).  Here ``forward`` of the model classes routes to
:class:`TrainStep` whenever gradients are required: ONE ``torch.autograd.Function`` that runs the forward with the same
kernels as inference (saving what the backward needs), and whose ``backward`` replays the layers in reverse with
  dgrad / wgrad GEMMs      ops.gemm(..., flags=GEMM_W_T / GEMM_A_T|GEMM_W_T)  -- tcgen05, MN-major operands, no transposes
  attention backward       ops.attention_bwd                                  -- two-pass tcgen05 flash backward
  LayerNorm + FDDT         ops.layernorm_fddt_bwd                             -- incl. gamma/beta and FDDT table gradients
  GELU'                    the dgrad GEMM's epilogue (EPI_DGELU_BF16)
  CTC / soft-label CE      ops.ctc_loss_fwd_bwd / ops.softlabel_ce_bwd
and returns fp32 gradients for exactly the parameters with ``requires_grad`` (so torch DDP / HF Trainer, gradient
clipping and any torch optimizer work unchanged; the gradient exchange is the one collective of the path).

Numerics follow the reference's bf16 autocast: bf16 GEMM / attention operands with fp32 accumulation, fp32 residual
stream, fp32 LayerNorm / softmax statistics / losses, fp32 gradients for fp32 master weights.

Three bridges into autograd (all return gradients for exactly the parameters that require grad):
  EncoderLogitsFn   DiCoWEncoder.forward(return_logits=True)  -> CTC logits with a grad_fn   (src/utils/trainers.py:76-103)
  CtcLossFn         DiCoWEncoder.get_loss(logits, labels)                                     (src/models/dicow/encoder.py:108-135)
  DiCoWTrainStepFn  DiCoWForConditionalGeneration.forward(labels=...) -> loss                 (modeling_dicow.py:248-354)

SE-DiCoW enrollment streams (src/models/dicow/layers.py:145-193) train through the same step: the training path keeps
the target and enrollment streams STACKED ([targets ; enrollments] on the batch axis, instead of the reference's
interleave -- no op on the path couples different batch entries other than the speaker communication block, which pairs
row b with row B + b), so both halves are contiguous GEMM / attention operands and gradient views.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops

_FDDT_ORDER = ("silence", "target", "non_target", "overlap")


def _ceil8(n: int) -> int:
    return -(-n // 8) * 8


_scalar_cache: Dict[tuple, torch.Tensor] = {}


def _dev_scalar(v: float, dev: torch.device) -> torch.Tensor:
    """device-resident fp32 scalar (alpha of the accumulate epilogue); cached so a step issues no H2D copies for it"""
    k = (float(v), dev.index)
    t = _scalar_cache.get(k)
    if t is None:
        t = _scalar_cache[k] = torch.tensor([v], dtype=torch.float32, device=dev)
    return t


# Set to a parallel.GradientExchange to all-reduce finished gradient groups on a communication stream while the backward
# is still running (the one exchange step of the path).  None: gradients are only handed to autograd (a DDP
# wrapper or parallel.allreduce_gradients can exchange them afterwards).
gradient_exchange = None


class _Grads:
    """fp32 gradient buffers keyed by parameter object; created on first use, only for parameters that require grad.
    ``reserve`` lays the gradients of a parameter group (one layer) out as views of ONE flat buffer so that the group is
    exchanged with a single collective and no packing copy (``flush``).
    A matrix whose row count is not a multiple of 8 (lm_head 51867, embed_tokens 51866) gets a row-padded buffer: the
    wgrad GEMM consumes dY MN-major and wants its M dimension a multiple of 8; the gradient handed to autograd is the
    [:rows] view."""

    def __init__(self, exchange=None):
        self.buf: Dict[int, torch.Tensor] = {}
        self.full: Dict[int, torch.Tensor] = {}
        self.params: Dict[int, torch.nn.Parameter] = {}
        self.exchange = exchange

    def reserve(self, params) -> Optional[torch.Tensor]:
        """allocate the (not yet existing) gradients of ``params`` contiguously; returns the flat buffer or None"""
        todo, seen, total = [], set(), 0
        for p in params:
            if p is None or not p.requires_grad or id(p) in self.buf or id(p) in seen:
                continue
            seen.add(id(p))
            shape = tuple(p.shape)
            if p.dim() >= 2 and shape[0] % 8:
                shape = (_ceil8(shape[0]),) + shape[1:]
            n = 1
            for v in shape:
                n *= v
            todo.append((p, shape, total, n))
            total += -(-n // 32) * 32  # 128-byte aligned segments (TMA / vector accesses on the views)
        if not todo:
            return None
        flat = torch.zeros(total, dtype=torch.float32, device=todo[0][0].device)
        for p, shape, off, n in todo:
            full = flat[off:off + n].view(shape)
            self.full[id(p)], self.buf[id(p)], self.params[id(p)] = full, full[:p.shape[0]] if p.dim() >= 1 else full, p
        return flat

    def flush(self, flat: Optional[torch.Tensor]) -> None:
        if flat is not None and self.exchange is not None:
            self.exchange.bucket_ready(flat)

    def finish(self, params) -> tuple:
        """the gradients in the order of ``params`` (None for frozen ones); drops this object's references so that
        autograd's AccumulateGrad can adopt the buffers instead of cloning them"""
        if self.exchange is not None:
            self.exchange.finish()
        out = tuple(self.buf.get(id(p)) if p.requires_grad else None for p in params)
        self.buf, self.full, self.params = {}, {}, {}
        return out

    def want(self, p: Optional[torch.nn.Parameter]) -> bool:
        return p is not None and p.requires_grad

    def get(self, p: torch.nn.Parameter) -> torch.Tensor:
        k = id(p)
        if k not in self.buf:
            if p.dim() >= 2 and p.shape[0] % 8:
                full = torch.zeros((_ceil8(p.shape[0]),) + tuple(p.shape[1:]), dtype=torch.float32, device=p.device)
                self.full[k], self.buf[k] = full, full[:p.shape[0]]
            else:
                self.buf[k] = self.full[k] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
            self.params[k] = p
        return self.buf[k]

    def get_padded(self, p: torch.nn.Parameter) -> torch.Tensor:
        self.get(p)
        return self.full[id(p)]


def _linear_backward(g: _Grads, dY: torch.Tensor, X: torch.Tensor, W_prep: torch.Tensor, weight: torch.nn.Parameter,
                     bias: Optional[torch.nn.Parameter], *, need_dx: bool = True, dx_epilogue: int = ops.EPI_BIAS_BF16,
                     aux: Optional[torch.Tensor] = None, scale: float = 1.0, dy_cols: Optional[slice] = None,
                     w_rows: Optional[slice] = None, dx_accum: Optional[torch.Tensor] = None,
                     bias_done: bool = False) -> Optional[torch.Tensor]:
    """Backward of Y = X W^T + b for bf16 operands: dX = dY W (bf16, optional gelu' epilogue; or accumulated into the
    fp32 ``dx_accum``), dW += scale * dY^T X, db += scale * colsum(dY).  ``dy_cols`` selects a column block of dY (fused
    projections), ``w_rows`` the matching rows of the prepared weight.  X [M, K], dY [M, N], W_prep [N, K] bf16."""
    dYs = dY if dy_cols is None else dY[:, dy_cols]
    Ws = W_prep if w_rows is None else W_prep[w_rows]
    dx = None
    if dx_accum is not None:
        ops.gemm(dYs, Ws, dx_accum, epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_W_T, lda=dY.stride(0), Mb=dY.shape[0],
                 splits=1)
    elif need_dx:
        dx = torch.empty(X.shape[0], Ws.shape[1], dtype=torch.bfloat16, device=X.device)
        ops.gemm(dYs, Ws, dx, epilogue=dx_epilogue, flags=ops.GEMM_W_T, aux=aux, lda=dY.stride(0), Mb=dY.shape[0])
    if g.want(weight):
        alpha = None if scale == 1.0 else _dev_scalar(scale, X.device)
        ops.gemm(dYs, X, g.get(weight).view(Ws.shape[0], -1), epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_A_T | ops.GEMM_W_T,
                 gate=alpha, lda=dY.stride(0), Mb=Ws.shape[0], K=X.shape[0], N=X.shape[1])
    if g.want(bias) and not bias_done:  # bias_done: the kernel that produced dY has added its column sums already
        ops.colsum(dYs, g.get(bias), alpha=scale)
    lora = getattr(weight, "_dicow_lora", None)
    if lora is not None and (g.want(lora.lora_A) or g.want(lora.lora_B)):
        _lora_backward(g, dYs, dY.stride(0), X, lora, scale)
    return dx


def _lora_backward(g: _Grads, dYs: torch.Tensor, ldy: int, X: torch.Tensor, lin, scale: float) -> None:
    """Gradients of a LoRA adapter Y = X (W + s B A)^T (the forward multiplies by the merged weight):
        dB += s * dY^T (X A^T)        dA += s * (dY B)^T X
    as four r-wide GEMMs on the tcgen05 kernel (r = 16: src/models/containers.py:71); ``scale`` is the factor already folded
    into the prepared weight (the q projections' hd^-0.5)."""
    A, B = lin.lora_A, lin.lora_B
    r, M = A.shape[0], X.shape[0]
    dev = X.device
    alpha = _dev_scalar(float(lin.lora_scale) * scale, dev)
    N_out = B.shape[0]
    if g.want(B):
        u = torch.empty(M, r, dtype=torch.bfloat16, device=dev)
        ops.gemm(X, ops.cast_bf16(A.detach().float()), u, epilogue=ops.EPI_BIAS_BF16)                  # u = X A^T
        ops.gemm(dYs, u, g.get_padded(B), epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_A_T | ops.GEMM_W_T, gate=alpha, lda=ldy,
                 Mb=_ceil8(N_out), K=M, N=r)                                                             # dB += s dY^T u
    if g.want(A):
        v = torch.empty(M, r, dtype=torch.bfloat16, device=dev)
        ops.gemm(dYs, ops.cast_bf16(B.detach().float()), v, epilogue=ops.EPI_BIAS_BF16, flags=ops.GEMM_W_T, lda=ldy, Mb=M,
                 K=N_out, N=r)                                                                           # v = dY B
        ops.gemm(v, X, g.get(A), epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_A_T | ops.GEMM_W_T, gate=alpha, lda=r, Mb=r, K=M,
                 N=X.shape[1])                                                                           # dA += s v^T X


class EncoderTape:
    """Activations saved by the training forward of DiCoWEncoder (one entry per layer)."""

    def __init__(self):
        self.layers: List[dict] = []
        self.stem: dict = {}
        self.final: dict = {}
        self.ctc: dict = {}


def _scb_forward_train(enc, e: dict, x: torch.Tensor, xb: torch.Tensor, B: int, T: int) -> dict:
    """SpeakerCommunicationBlock (src/models/dicow/layers.py:145-170) on the stacked streams: x fp32 [2B*T, d] (target rows
    [:B*T] updated in place), xb its bf16 copy.  Returns the activations the backward needs."""
    cfg = enc.config
    d, H, ffn = cfg.d_model, cfg.encoder_attention_heads, e["w1"].shape[0]
    dev = x.device
    BT = B * T
    xq, xe = xb[:BT], xb[BT:]
    q = torch.empty(BT, d, dtype=torch.bfloat16, device=dev)
    kv = torch.empty(BT, 2 * d, dtype=torch.bfloat16, device=dev)
    ops.gemm(xq, e["wq"], q, epilogue=ops.EPI_BIAS_BF16, bias=e["bq"])
    ops.gemm(xe, e["wkv"], kv, epilogue=ops.EPI_BIAS_BF16, bias=e["bkv"])
    ctx = torch.empty(BT, d, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, T, dtype=torch.float32, device=dev)
    ops.attention(q, kv, kv[:, d:], ctx, B=B, H=H, Tq=T, Tk=T, q_row_stride=d, q_batch_stride=T * d, kv_row_stride=2 * d,
                  kv_batch_stride=T * 2 * d, o_row_stride=d, o_batch_stride=T * d, lse=lse)
    ao = torch.empty(BT, d, dtype=torch.bfloat16, device=dev)
    ops.gemm(ctx, e["wo"], ao, epilogue=ops.EPI_BIAS_BF16, bias=e["bo"])
    hdn = torch.empty(BT, ffn, dtype=torch.bfloat16, device=dev)
    pre = torch.empty(BT, ffn, dtype=torch.bfloat16, device=dev)
    # ffn.0 over cat([attn_out, q_stream]) as a two-source contraction (no concat buffer)
    ops.gemm(ao, e["w1"], hdn, epilogue=ops.EPI_GELU_SAVE_BF16, bias=e["b1"], Mb=BT, K=2 * d, lda=d, A2=xq, lda2=d, K1=d,
             aux=pre)
    xt = x[:BT]
    ops.gemm(hdn, e["w2"], xt, epilogue=ops.EPI_RESIDUAL_F32, bias=e["b2"], resid=xt, ldr=d, gate=e["gate"])
    return {"xb": xb, "q": q, "kv": kv, "ctx": ctx, "lse": lse, "ao": ao, "pre": pre, "hdn": hdn}


def encoder_forward_train(enc, input_features: torch.Tensor, stno_mask: Optional[torch.Tensor], tape: EncoderTape,
                          enrollments: Optional[dict] = None):
    """DiCoWEncoder.forward (src/models/dicow/encoder.py:140-246) in training mode: same kernels as inference, keeping
    the per-layer inputs the backward needs.  Returns (last_hidden fp32 [B, T, d], bf16 copy [B*T, d]).
    With ``enrollments`` the first ``scb_layers`` layers run on the stacked [targets ; enrollments] streams."""
    cfg = enc.config
    n_scb = cfg.scb_layers if (cfg.use_enrollments and cfg.scb_layers and enrollments is not None) else 0
    if n_scb:  # encoder.py:152-154 (stacked instead of interleaved, see the module docstring)
        input_features = torch.cat((input_features, enrollments["input_features"].to(input_features.device)), dim=0)
        if stno_mask is not None:
            stno_mask = torch.cat((stno_mask, enrollments["stno_mask"].to(stno_mask.device)), dim=0)
    w = enc.prepare()
    dev = input_features.device
    d, F = cfg.d_model, input_features.shape[-1]
    B, T = input_features.shape[0], F // 2
    if F != enc.get_max_len():
        raise ValueError(f"Whisper expects the mel input features to be of length {enc.get_max_len()}, but found {F}. "
                         f"Make sure to pad the input mel features to {enc.get_max_len()}.")
    feats = input_features.float().contiguous()
    stno = stno_mask.to(device=dev, dtype=torch.float32).contiguous() if cfg.use_fddt else None
    C = cfg.num_mel_bins
    a0 = torch.empty(B, F + 2, C, dtype=torch.bfloat16, device=dev)
    ops.features_to_channels_last(feats, a0)
    a1 = torch.empty(B, F + 2, d, dtype=torch.bfloat16, device=dev)
    ops.zero_pad_rows(a1)
    pre1 = torch.zeros(B, F + 2, d, dtype=torch.bfloat16, device=dev)
    ops.gemm(a0, w["conv1_w"], a1[:, 1:], epilogue=ops.EPI_GELU_SAVE_BF16, bias=w["conv1_b"], nb=B, Mb=F, K=3 * C, lda=C,
             a_batch_stride=(F + 2) * C, ldo=d, out_batch_stride=(F + 2) * d, aux=pre1[:, 1:])
    x = torch.empty(B, T, d, dtype=torch.float32, device=dev)
    if stno is not None and cfg.use_pre_pos_fddt:
        stno0 = stno
    else:
        stno0 = torch.zeros(B, 4, T, dtype=torch.float32, device=dev)
        stno0[:, 0] = 1.0
    fw0, fb0 = w["fddt0"]
    full0 = bool(w.get("fddt_full")) and w.get("fddt0_full") is not None
    if full0:  # full-matrix initial FDDT (encoder.py:173-179 with CustomLinear): see DiCoWEncoder._forward_hidden
        g2 = torch.empty(B * T, d, dtype=torch.bfloat16, device=dev)
        ops.gemm(a1, w["conv2_w"], g2, epilogue=ops.EPI_BIAS_GELU_BF16, bias=w["conv2_b"], nb=B, Mb=T, K=3 * d, lda=2 * d,
                 a_batch_stride=(F + 2) * d, ldo=d, out_batch_stride=T * d)
        W4, b4 = w["fddt0_full"]
        y = torch.empty(B * T, 4 * d, dtype=torch.bfloat16, device=dev)
        ops.gemm(g2, W4, y, epilogue=ops.EPI_BIAS_BF16, bias=b4)
        ops.fddt_full_combine(y, stno0, x, T=T, pos=w["pos"])
        del g2, y
    else:
        ops.gemm(a1, w["conv2_w"], x, epilogue=ops.EPI_GELU_FDDT_POS_F32, bias=w["conv2_b"], nb=B, Mb=T, K=3 * d, lda=2 * d,
                 a_batch_stride=(F + 2) * d, ldo=d, out_batch_stride=T * d, stno=stno0, stno_batch_stride=4 * T, fddt_w=fw0,
                 fddt_b=fb0, pos=w["pos"])
    tape.stem = {"a0": a0, "a1": a1, "pre1": pre1, "stno0": stno0, "B": B, "T": T, "F": F, "full": full0}
    rows = B * T
    ffn = cfg.encoder_ffn_dim
    H = cfg.encoder_attention_heads
    d1 = d2 = None
    x = x.view(rows, d)
    if n_scb > len(w["layers"]):
        raise ValueError("scb_layers exceeds the number of encoder layers")
    for i, e in enumerate(w["layers"]):
        fd = w["fddt"][i] if (cfg.use_fddt and i < len(w["fddt"])) else None
        x_pre = x_in = x  # the stream before this layer's pending deltas / FDDT is an input of the backward:
        x = torch.empty_like(x_pre)  # the first kernel of the layer reads x_in and writes the updated stream to x (no copy)
        scb = None
        rows_in, stno_in = rows, stno
        full = None
        if fd is not None and w.get("fddt_full"):
            # full-matrix FDDT (layers.py:7-47, FDDT.py:52-62): fold the pending deltas, project the bf16 stream with the
            # stacked [4 d, d] class transforms, mask-weighted sum back into the fp32 stream.  For the rest of the layer
            # (and its backward) the transformed stream is the layer input: no table, no pending deltas.
            xb_in = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            ops.fddt_layernorm(x_in, x_out_bf16=xb_in, delta1=d1, delta2=d2, store_x=True, x_out=x)
            y = torch.empty(rows, 4 * d, dtype=torch.bfloat16, device=dev)
            ops.gemm(xb_in, fd[0], y, epilogue=ops.EPI_BIAS_BF16, bias=fd[1])
            ops.fddt_full_combine(y, stno, x, T=T)
            del y
            full = {"xb": xb_in, "W4": fd[0]}
            x_pre, x_in, fd, d1, d2 = x, x, None, None, None
        if i < n_scb:  # encoder.py:205-213: FDDT, speaker communication block, (last SCB layer) drop the enrollment stream
            xb = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            ops.fddt_layernorm(x_in, T=T, stno=stno if fd else None, fddt_w=fd[0] if fd else None,
                               fddt_b=fd[1] if fd else None, x_out_bf16=xb, delta1=d1, delta2=d2, store_x=True,
                               x_out=None if x_in is x else x)
            scb = _scb_forward_train(enc, w["scb"][i], x, xb, B // 2, T)
            if i == n_scb - 1:
                B //= 2
                rows = B * T
                x = x[:rows]
                stno = stno[:B] if stno is not None else None
            ln1 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            ops.fddt_layernorm(x, gamma=e["ln1_g"], beta=e["ln1_b"], ln_out_bf16=ln1, store_x=False)
        else:
            ln1 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            ops.fddt_layernorm(x_in, T=T, stno=stno if fd else None, fddt_w=fd[0] if fd else None,
                               fddt_b=fd[1] if fd else None, gamma=e["ln1_g"], beta=e["ln1_b"], ln_out_bf16=ln1, delta1=d1,
                               delta2=d2, store_x=True, x_out=None if x_in is x else x)
        qkv = torch.empty(rows, 3 * d, dtype=torch.bfloat16, device=dev)
        ops.gemm(ln1, e["wqkv"], qkv, epilogue=ops.EPI_BIAS_BF16, bias=e["bqkv"])
        ctx = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        lse = torch.empty(B, H, T, dtype=torch.float32, device=dev)
        ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], ctx, B=B, H=H, Tq=T, Tk=T, q_row_stride=3 * d,
                      q_batch_stride=T * 3 * d, kv_row_stride=3 * d, kv_batch_stride=T * 3 * d, o_row_stride=d,
                      o_batch_stride=T * d, lse=lse)
        d1n = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ops.gemm(ctx, e["wo"], d1n, epilogue=ops.EPI_BIAS_BF16, bias=e["bo"])
        ln2 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ops.fddt_layernorm(x, gamma=e["ln2_g"], beta=e["ln2_b"], ln_out_bf16=ln2, delta1=d1n, store_x=False)
        hdn = torch.empty(rows, ffn, dtype=torch.bfloat16, device=dev)
        pre = torch.empty(rows, ffn, dtype=torch.bfloat16, device=dev)
        ops.gemm(ln2, e["w1"], hdn, epilogue=ops.EPI_GELU_SAVE_BF16, bias=e["b1"], aux=pre)
        d2n = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ops.gemm(hdn, e["w2"], d2n, epilogue=ops.EPI_BIAS_BF16, bias=e["b2"])
        tape.layers.append({"x_pre": x_pre, "d1_in": d1, "d2_in": d2, "x_post": x, "ln1": ln1, "qkv": qkv, "ctx": ctx,
                            "lse": lse, "d1": d1n, "ln2": ln2, "pre": pre, "hdn": hdn, "d2": d2n, "fd": fd, "scb": scb,
                            "B": B, "rows_in": rows_in, "stno_in": stno_in, "full": full})
        d1, d2 = d1n, d2n
    out = torch.empty(B, T, d, dtype=torch.float32, device=dev)
    out_bf16 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
    ops.fddt_layernorm(x, gamma=w["lnf_g"], beta=w["lnf_b"], ln_out_f32=out, ln_out_bf16=out_bf16, delta1=d1, delta2=d2,
                       store_x=False)
    tape.final = {"x": x, "d1": d1, "d2": d2, "stno": stno, "B": B, "T": T}
    return out, out_bf16


def ctc_head_forward_train(enc, hidden_bf16: torch.Tensor, B: int, T: int, tape: EncoderTape) -> torch.Tensor:
    """possibly_update_last_hidden_states + lm_head (encoder.py:87-106,236) keeping the backward's inputs."""
    cfg = enc.config
    w = enc.prepare()
    d, H = cfg.d_model, cfg.encoder_attention_heads
    dev = hidden_bf16.device
    if ("ctc_attn" not in w and "ctc_layer" not in w) or "sub1" not in w:
        raise NotImplementedError("training the CTC head needs additional_self_attention_layer (or additional_layer) + "
                                  "pre_ctc_sub_sample (the recipes' configuration)")
    rows = B * T
    full = "ctc_layer" in w  # a whole pre-LN encoder layer instead of the bare self-attention (encoder.py:88-93)
    e = w["ctc_layer"] if full else w["ctc_attn"]
    extra = {}
    attn_in = hidden_bf16
    if full:
        x = hidden_bf16.float()  # the layer's fp32 residual stream
        attn_in = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ops.fddt_layernorm(x, gamma=e["ln1_g"], beta=e["ln1_b"], ln_out_bf16=attn_in, store_x=False)
    qkv = torch.empty(rows, 3 * d, dtype=torch.bfloat16, device=dev)
    ops.gemm(attn_in, e["wqkv"], qkv, epilogue=ops.EPI_BIAS_BF16, bias=e["bqkv"])
    ctx = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, T, dtype=torch.float32, device=dev)
    ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], ctx, B=B, H=H, Tq=T, Tk=T, q_row_stride=3 * d, q_batch_stride=T * 3 * d,
                  kv_row_stride=3 * d, kv_batch_stride=T * 3 * d, o_row_stride=d, o_batch_stride=T * d, lse=lse)
    buf = torch.empty(B, T + 2, d, dtype=torch.bfloat16, device=dev)
    ops.zero_pad_rows(buf)
    if full:
        ffn = e["w1"].shape[0]
        d1 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ops.gemm(ctx, e["wo"], d1, epilogue=ops.EPI_BIAS_BF16, bias=e["bo"])
        ln2 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ops.fddt_layernorm(x, gamma=e["ln2_g"], beta=e["ln2_b"], ln_out_bf16=ln2, delta1=d1, store_x=False)
        hdn = torch.empty(rows, ffn, dtype=torch.bfloat16, device=dev)
        pre = torch.empty(rows, ffn, dtype=torch.bfloat16, device=dev)
        ops.gemm(ln2, e["w1"], hdn, epilogue=ops.EPI_GELU_SAVE_BF16, bias=e["b1"], aux=pre)
        d2 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ops.gemm(hdn, e["w2"], d2, epilogue=ops.EPI_BIAS_BF16, bias=e["b2"])
        out_bf16 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ops.fddt_layernorm(x, x_out_bf16=out_bf16, delta1=d1, delta2=d2, store_x=False)  # x + d1 + d2 -> bf16
        buf[:, 1:T + 1] = out_bf16.view(B, T, d)
        extra = {"x": x, "ln1": attn_in, "d1": d1, "ln2": ln2, "hdn": hdn, "pre": pre}
    else:
        ops.gemm(ctx, e["wo"], buf[:, 1:], epilogue=ops.EPI_BIAS_BF16, bias=e["bo"], nb=B, Mb=T, lda=d, a_batch_stride=T * d,
                 ldo=d, out_batch_stride=(T + 2) * d)
    T1 = (T + 2 - 3) // 2 + 1
    buf1 = torch.empty(B, T1 + 2, d, dtype=torch.bfloat16, device=dev)
    ops.zero_pad_rows(buf1)
    ops.gemm(buf, w["sub1"], buf1[:, 1:], epilogue=ops.EPI_BIAS_BF16, nb=B, Mb=T1, K=3 * d, lda=2 * d,
             a_batch_stride=(T + 2) * d, ldo=d, out_batch_stride=(T1 + 2) * d)
    T2 = (T1 + 2 - 3) // 2 + 1
    neck = torch.empty(B, T2, d, dtype=torch.bfloat16, device=dev)
    ops.gemm(buf1, w["sub2"], neck, epilogue=ops.EPI_BIAS_BF16, nb=B, Mb=T2, K=3 * d, lda=2 * d,
             a_batch_stride=(T1 + 2) * d, ldo=d, out_batch_stride=T2 * d)
    V1 = w["lm_head"].shape[0]
    # rows padded to a multiple of 4 floats (V + 1 = 51 867 is odd): the GEMM epilogue stores 16-byte vectors instead of 4-byte
    # scalars (2.65 -> 0.9 ms at B = 16); the CTC kernels take the row stride
    ldl = -(-V1 // 4) * 4
    logits = torch.empty(B, T2, ldl, dtype=torch.float32, device=dev)[..., :V1]
    ops.gemm(neck.view(B * T2, d), w["lm_head"], logits.as_strided((B * T2, V1), (ldl, 1)), epilogue=ops.EPI_BIAS_F32)
    tape.ctc = {"hidden": hidden_bf16, "qkv": qkv, "ctx": ctx, "lse": lse, "buf": buf, "buf1": buf1, "neck": neck,
                "T1": T1, "T2": T2, "B": B, "T": T, **extra}
    return logits


def _conv_backward(g: _Grads, dY: torch.Tensor, Xpad: torch.Tensor, W_prep: torch.Tensor, weight: torch.nn.Parameter,
                   bias: Optional[torch.nn.Parameter], *, B: int, T_in: int, T_out: int, C_in: int, stride: int,
                   need_dx: bool, dx_padded: bool = False, bias_src: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Backward of the implicit-GEMM Conv1d(k=3, p=1): dY bf16 [B, T_out, Cout] (each dY[b] contiguous), Xpad the
    zero-padded channels-last input [B, T_in + 2, C_in].  Returns dX bf16 (col2im of dY W) when requested: [B, T_in, C_in],
    or with ``dx_padded`` the zero-padded [B, T_in + 2, C_in] layout of the forward buffer."""
    dev = dY.device
    Cout = W_prep.shape[0]
    dx = None
    if need_dx:
        dYf = dY.reshape(B * T_out, Cout)
        dcol = torch.empty(B * T_out, 3 * C_in, dtype=torch.bfloat16, device=dev)
        ops.gemm(dYf, W_prep, dcol, epilogue=ops.EPI_BIAS_BF16, flags=ops.GEMM_W_T)
        if dx_padded:
            dx = torch.zeros(B, T_in + 2, C_in, dtype=torch.bfloat16, device=dev)
            ops.conv1d_col2im(dcol, dx[:, 1:], B=B, T=T_in, T_out=T_out, C_in=C_in, stride=stride,
                              dx_batch_stride=(T_in + 2) * C_in, dx_row_stride=C_in)
        else:
            dx = torch.empty(B, T_in, C_in, dtype=torch.bfloat16, device=dev)
            ops.conv1d_col2im(dcol, dx, B=B, T=T_in, T_out=T_out, C_in=C_in, stride=stride, dx_batch_stride=T_in * C_in,
                              dx_row_stride=C_in)
    if g.want(weight):
        # dW[Cout, 3 C_in] += dY_b^T im2col(X_b): the im2col view is the overlapping-row operand (row stride stride*C_in)
        gw = g.get(weight)  # parameter layout [Cout, C_in, 3]; accumulate tap-major then fold back at the end
        acc = g.__dict__.setdefault("_conv_acc", {})
        key = id(weight)
        if key not in acc:
            acc[key] = (torch.zeros(Cout, 3 * C_in, dtype=torch.float32, device=dev), gw, C_in)
        for b in range(B):
            ops.gemm(dY[b], Xpad[b], acc[key][0], epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_A_T | ops.GEMM_W_T,
                     lda=Cout, Mb=Cout, K=T_out, N=3 * C_in, ldw=stride * C_in)
    if g.want(bias):
        ops.colsum(bias_src if bias_src is not None else dY.reshape(B * T_out, Cout), g.get(bias))
    return dx


def _fold_conv_grads(g: _Grads) -> None:
    for acc, gw, C_in in g.__dict__.get("_conv_acc", {}).values():
        gw.add_(acc.view(gw.shape[0], 3, C_in).permute(0, 2, 1))
    g.__dict__["_conv_acc"] = {}


def _attention_params_backward(g: _Grads, att_mod, e: dict, dqkv: torch.Tensor, X: torch.Tensor, d: int, need_dx: bool,
                               dx_accum: Optional[torch.Tensor] = None):
    """q/k/v projections of a fused [3d, d] prepared weight (q rows carry the folded 1/8 scale).  dX is returned as
    bf16, or accumulated into the fp32 ``dx_accum`` when given."""
    dx = None
    if dx_accum is not None:
        ops.gemm(dqkv, e["wqkv"], dx_accum, epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_W_T, splits=1)
    elif need_dx:
        dx = torch.empty(X.shape[0], d, dtype=torch.bfloat16, device=X.device)
        ops.gemm(dqkv, e["wqkv"], dx, epilogue=ops.EPI_BIAS_BF16, flags=ops.GEMM_W_T)
    sc = 64 ** -0.5
    _linear_backward(g, dqkv, X, e["wqkv"], att_mod.q_proj.weight, att_mod.q_proj.bias, need_dx=False, scale=sc,
                     dy_cols=slice(0, d), w_rows=slice(0, d))
    _linear_backward(g, dqkv, X, e["wqkv"], att_mod.k_proj.weight, None, need_dx=False, dy_cols=slice(d, 2 * d),
                     w_rows=slice(d, 2 * d))
    _linear_backward(g, dqkv, X, e["wqkv"], att_mod.v_proj.weight, att_mod.v_proj.bias, need_dx=False,
                     dy_cols=slice(2 * d, 3 * d), w_rows=slice(2 * d, 3 * d))
    return dx


def ctc_head_backward(enc, g: _Grads, dlogits: torch.Tensor, tape: EncoderTape, need_dhidden: bool,
                      dhidden_accum: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """dlogits bf16 [B, T2, ldd] (ldd = ceil8(V + 1), padding columns zero) -> parameter gradients of the head; returns
    dL/d(hidden) bf16, or accumulates it into the fp32 ``dhidden_accum``."""
    cfg = enc.config
    w = enc.prepare()
    c = tape.ctc
    B, T, T1, T2, d, H = c["B"], c["T"], c["T1"], c["T2"], cfg.d_model, cfg.encoder_attention_heads
    V1 = w["lm_head"].shape[0]
    dl = dlogits.view(B * T2, -1)
    ldd = dl.shape[1]
    full = "ctc_layer" in w
    front = enc.additional_layer if full else enc.additional_self_attention_layer
    flat = g.reserve([enc.lm_head.weight, enc.subsample_conv1.weight, enc.subsample_conv2.weight] + list(front.parameters()))
    # lm_head (no bias): dneck = dlogits W ; dW += dlogits^T neck
    dneck = torch.empty(B * T2, d, dtype=torch.bfloat16, device=dl.device)
    ops.gemm(dl, w["lm_head"], dneck, epilogue=ops.EPI_BIAS_BF16, flags=ops.GEMM_W_T, K=V1, lda=ldd, Mb=B * T2)
    if g.want(enc.lm_head.weight):
        ops.gemm(dl, c["neck"].view(B * T2, d), g.get_padded(enc.lm_head.weight), epilogue=ops.EPI_ACCUM_F32,
                 flags=ops.GEMM_A_T | ops.GEMM_W_T, lda=ldd, Mb=ldd, K=B * T2, N=d)
    dbuf1 = _conv_backward(g, dneck.view(B, T2, d), c["buf1"], w["sub2"], enc.subsample_conv2.weight, None, B=B, T_in=T1,
                           T_out=T2, C_in=d, stride=2, need_dx=True)
    dbuf = _conv_backward(g, dbuf1, c["buf"], w["sub1"], enc.subsample_conv1.weight, None, B=B, T_in=T, T_out=T1, C_in=d,
                          stride=2, need_dx=True)
    dbuf_f = dbuf.view(B * T, d)
    if full:
        dh = _ctc_layer_backward(enc, g, dbuf_f, c, w["ctc_layer"], B, T, need_dhidden, dhidden_accum)
        _fold_conv_grads(g)
        g.flush(flat)
        return dh
    att, e = enc.additional_self_attention_layer, w["ctc_attn"]
    dctx = _linear_backward(g, dbuf_f, c["ctx"], e["wo"], att.out_proj.weight, att.out_proj.bias)
    dqkv = torch.empty(B * T, 3 * d, dtype=torch.bfloat16, device=dl.device)
    qkv = c["qkv"]
    ops.attention_bwd(qkv, qkv[:, d:], qkv[:, 2 * d:], c["ctx"], dctx, c["lse"], dqkv, dqkv[:, d:], dqkv[:, 2 * d:], B=B,
                      H=H, Tq=T, Tk=T, q_row_stride=3 * d, q_batch_stride=T * 3 * d, kv_row_stride=3 * d,
                      kv_batch_stride=T * 3 * d, o_row_stride=d, o_batch_stride=T * d, dq_row_stride=3 * d,
                      dq_batch_stride=T * 3 * d, dkv_row_stride=3 * d, dkv_batch_stride=T * 3 * d)
    dh = _attention_params_backward(g, att, e, dqkv, c["hidden"], d, need_dhidden,
                                    dx_accum=dhidden_accum if need_dhidden else None)
    _fold_conv_grads(g)
    g.flush(flat)
    return dh


def _ctc_layer_backward(enc, g: _Grads, dout: torch.Tensor, c: dict, e: dict, B: int, T: int, need_dhidden: bool,
                        dhidden_accum: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """Backward of the whole encoder layer in front of the CTC head (``additional_layer=True``, encoder.py:88-93; the layer is
    HF WhisperEncoderLayer, HF:modeling_whisper.py:380-430).  dout bf16 [B*T, d] = gradient of the layer output x + d1 + d2."""
    cfg = enc.config
    lyr = enc.additional_layer
    d, H = cfg.d_model, cfg.encoder_attention_heads
    rows, dev = B * T, dout.device
    G = dout.float()
    dpre = _linear_backward(g, dout, c["hdn"], e["w2"], lyr.fc2.weight, lyr.fc2.bias, dx_epilogue=ops.EPI_DGELU_BF16,
                            aux=c["pre"])
    dln2 = _linear_backward(g, dpre, c["ln2"], e["w1"], lyr.fc1.weight, lyr.fc1.bias)
    G2 = torch.empty(rows, d, dtype=torch.float32, device=dev)
    G2b = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
    n2, n1 = lyr.final_layer_norm, lyr.self_attn_layer_norm
    ops.layernorm_fddt_bwd(c["x"], G2, dy=dln2, g_in=G, gamma=e["ln2_g"], delta1=c["d1"], g_out_bf16=G2b,
                           dgamma=g.get(n2.weight) if g.want(n2.weight) else None,
                           dbeta=g.get(n2.bias) if g.want(n2.bias) else None)
    dctx = _linear_backward(g, G2b, c["ctx"], e["wo"], lyr.self_attn.out_proj.weight, lyr.self_attn.out_proj.bias)
    dqkv = torch.empty(rows, 3 * d, dtype=torch.bfloat16, device=dev)
    qkv = c["qkv"]
    ops.attention_bwd(qkv, qkv[:, d:], qkv[:, 2 * d:], c["ctx"], dctx, c["lse"], dqkv, dqkv[:, d:], dqkv[:, 2 * d:], B=B,
                      H=H, Tq=T, Tk=T, q_row_stride=3 * d, q_batch_stride=T * 3 * d, kv_row_stride=3 * d,
                      kv_batch_stride=T * 3 * d, o_row_stride=d, o_batch_stride=T * d, dq_row_stride=3 * d,
                      dq_batch_stride=T * 3 * d, dkv_row_stride=3 * d, dkv_batch_stride=T * 3 * d)
    dln1 = _attention_params_backward(g, lyr.self_attn, e, dqkv, c["ln1"], d, True)
    want_g1 = g.want(n1.weight)
    if not need_dhidden and not want_g1 and not g.want(n1.bias):
        return None
    Gh = torch.empty(rows, d, dtype=torch.float32, device=dev)
    Ghb = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
    ops.layernorm_fddt_bwd(c["x"], Gh, dy=dln1, g_in=G2, gamma=e["ln1_g"], g_out_bf16=Ghb,
                           dgamma=g.get(n1.weight) if want_g1 else None, dbeta=g.get(n1.bias) if g.want(n1.bias) else None)
    if not need_dhidden:
        return None
    if dhidden_accum is not None:
        dhidden_accum.add_(Gh)
        return None
    return Ghb


def _scb_backward(enc, g: _Grads, cae, e: dict, a: dict, Gs: torch.Tensor, B: int, T: int) -> None:
    """Backward of _scb_forward_train.  Gs fp32 [2B*T, d]: on entry the gradient of the block's OUTPUT streams (targets
    [:B*T], enrollments [B*T:]); on exit the gradient of its INPUT streams (the FDDT output), accumulated in place:
      targets      += d(ffn.0)/d(q half of the concat) + dq Wq           enrollments += dkv Wkv
    (the target rows pass straight through the residual, src/models/dicow/layers.py:168)."""
    cfg = enc.config
    d, H, ffn = cfg.d_model, cfg.encoder_attention_heads, e["w1"].shape[0]
    dev = Gs.device
    BT = B * T
    Gt, Ge = Gs[:BT], Gs[BT:]
    xq, xe = a["xb"][:BT], a["xb"][BT:]
    lin1, lin2 = cae.ffn[0], cae.ffn[3]
    # gate: out = q + tanh(gate) * upd, upd = ffn.3(hdn) recomputed (one GEMM) instead of stored
    upd = torch.empty(BT, d, dtype=torch.bfloat16, device=dev)
    ops.gemm(a["hdn"], e["w2"], upd, epilogue=ops.EPI_BIAS_BF16, bias=e["b2"])
    gate = cae.cross_gate.gate
    dupd = ops.gate_bwd(Gt, upd, e["gate"], g.get(gate).view(-1) if g.want(gate) else None)
    del upd
    dpre = _linear_backward(g, dupd, a["hdn"], e["w2"], lin2.weight, lin2.bias, dx_epilogue=ops.EPI_DGELU_BF16, aux=a["pre"])
    # ffn.0 over the concat [attn_out | q stream]: dgrad per half (the q half goes straight into the fp32 stream
    # gradient), wgrad per column half of the [ffn, 2d] weight
    w1 = e["w1"]
    dao = torch.empty(BT, d, dtype=torch.bfloat16, device=dev)
    ops.gemm(dpre, w1[:, :d], dao, epilogue=ops.EPI_BIAS_BF16, flags=ops.GEMM_W_T, K=ffn, N=d)
    ops.gemm(dpre, w1[:, d:], Gt, epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_W_T, K=ffn, N=d, splits=1)
    if g.want(lin1.weight):
        gw1 = g.get(lin1.weight)
        for half, X in ((gw1[:, :d], a["ao"]), (gw1[:, d:], xq)):
            ops.gemm(dpre, X, half, epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_A_T | ops.GEMM_W_T, lda=dpre.stride(0), Mb=ffn,
                     K=BT, N=d, ldo=2 * d)
    if g.want(lin1.bias):
        ops.colsum(dpre, g.get(lin1.bias))
    del dpre
    att = cae.cross_attn
    dctx = _linear_backward(g, dao, a["ctx"], e["wo"], att.out_proj.weight, att.out_proj.bias)
    dq = torch.empty(BT, d, dtype=torch.bfloat16, device=dev)
    dkv = torch.empty(BT, 2 * d, dtype=torch.bfloat16, device=dev)
    q, kv = a["q"], a["kv"]
    ops.attention_bwd(q, kv, kv[:, d:], a["ctx"], dctx, a["lse"], dq, dkv, dkv[:, d:], B=B, H=H, Tq=T, Tk=T, q_row_stride=d,
                      q_batch_stride=T * d, kv_row_stride=2 * d, kv_batch_stride=T * 2 * d, o_row_stride=d,
                      o_batch_stride=T * d, dq_row_stride=d, dq_batch_stride=T * d, dkv_row_stride=2 * d,
                      dkv_batch_stride=T * 2 * d)
    sc = 64 ** -0.5  # folded into the prepared Wq / bq (modeling._prep_attention)
    _linear_backward(g, dq, xq, e["wq"], att.q_proj.weight, att.q_proj.bias, scale=sc, dx_accum=Gt)
    ops.gemm(dkv, e["wkv"], Ge, epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_W_T, splits=1)
    _linear_backward(g, dkv, xe, e["wkv"], att.k_proj.weight, None, need_dx=False, dy_cols=slice(0, d), w_rows=slice(0, d))
    _linear_backward(g, dkv, xe, e["wkv"], att.v_proj.weight, att.v_proj.bias, need_dx=False, dy_cols=slice(d, 2 * d),
                     w_rows=slice(d, 2 * d))


def encoder_backward(enc, g: _Grads, d_hidden_bf16: torch.Tensor, tape: EncoderTape) -> None:
    """Backward of encoder_forward_train given dL/d(last_hidden_state) as bf16 [B*T, d]."""
    cfg = enc.config
    w = enc.prepare()
    f = tape.final
    B, T, d, H = f["B"], f["T"], cfg.d_model, cfg.encoder_attention_heads
    rows = B * T
    dev = d_hidden_bf16.device
    stno = f["stno"]
    G = torch.empty(rows, d, dtype=torch.float32, device=dev)
    Gb = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
    ln = enc.layer_norm
    n_layers = len(w["layers"])

    def group(i):  # one exchange bucket per layer: its own parameters, its FDDT tables, its speaker communication block
        ps = list(enc.layers[i].parameters())  # (+ the final LayerNorm with the last layer)
        if tape.layers[i]["fd"] is not None or tape.layers[i]["full"] is not None:
            ps += list(enc.fddts[i].parameters())
        if tape.layers[i]["scb"] is not None:
            ps += list(enc.ca_enrolls[i].parameters())
        return ps + ([ln.weight, ln.bias] if i == n_layers - 1 else [])

    flat = g.reserve(group(n_layers - 1)) if n_layers else g.reserve([ln.weight, ln.bias])

    def fc2_bias_grad(j):  # Gb is dY of layer j's fc2: the kernel that writes Gb also adds its column sums to the bias gradient
        if j < 0 or not g.want(enc.layers[j].fc2.bias):
            return None
        return g.get(enc.layers[j].fc2.bias)

    gs = fc2_bias_grad(n_layers - 1)
    ops.layernorm_fddt_bwd(f["x"], G, dy=d_hidden_bf16, gamma=w["lnf_g"], delta1=f["d1"], delta2=f["d2"], g_out_bf16=Gb,
                           dgamma=g.get(ln.weight) if g.want(ln.weight) else None,
                           dbeta=g.get(ln.bias) if g.want(ln.bias) else None, g_colsum=gs)
    fc2_bias_done = gs is not None
    flat_next = None
    for i in range(n_layers - 1, -1, -1):
        e, s, lyr = w["layers"][i], tape.layers[i], enc.layers[i]
        B = s["B"]  # streams in this layer's attention / MLP blocks (targets only after the last SCB layer)
        rows = B * T
        if i != n_layers - 1:
            flat = flat_next if flat_next is not None else g.reserve(group(i))
        # fc2 / fc1 (G is the gradient of x_post + d1 + d2, hence of d2 = fc2(...) as well)
        dpre = _linear_backward(g, Gb, s["hdn"], e["w2"], lyr.fc2.weight, lyr.fc2.bias, dx_epilogue=ops.EPI_DGELU_BF16,
                                aux=s["pre"], bias_done=fc2_bias_done)
        dln2 = _linear_backward(g, dpre, s["ln2"], e["w1"], lyr.fc1.weight, lyr.fc1.bias)
        G2 = torch.empty(rows, d, dtype=torch.float32, device=dev)
        G2b = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        n2 = lyr.final_layer_norm
        ob = lyr.self_attn.out_proj.bias  # G2b is dY of out_proj: its column sums (the bias gradient) come with it
        ops.layernorm_fddt_bwd(s["x_post"], G2, dy=dln2, g_in=G, gamma=e["ln2_g"], delta1=s["d1"], g_out_bf16=G2b,
                               dgamma=g.get(n2.weight) if g.want(n2.weight) else None,
                               dbeta=g.get(n2.bias) if g.want(n2.bias) else None, g_colsum=g.get(ob) if g.want(ob) else None)
        # attention block
        dctx = _linear_backward(g, G2b, s["ctx"], e["wo"], lyr.self_attn.out_proj.weight, ob, bias_done=True)
        dqkv = torch.empty(rows, 3 * d, dtype=torch.bfloat16, device=dev)
        qkv = s["qkv"]
        ops.attention_bwd(qkv, qkv[:, d:], qkv[:, 2 * d:], s["ctx"], dctx, s["lse"], dqkv, dqkv[:, d:], dqkv[:, 2 * d:],
                          B=B, H=H, Tq=T, Tk=T, q_row_stride=3 * d, q_batch_stride=T * 3 * d, kv_row_stride=3 * d,
                          kv_batch_stride=T * 3 * d, o_row_stride=d, o_batch_stride=T * d, dq_row_stride=3 * d,
                          dq_batch_stride=T * 3 * d, dkv_row_stride=3 * d, dkv_batch_stride=T * 3 * d)
        dln1 = _attention_params_backward(g, lyr.self_attn, e, dqkv, s["ln1"], d, True)
        # LayerNorm 1 + FDDT of this layer; the result is the gradient of (x_pre + d1_in + d2_in)
        n1 = lyr.self_attn_layer_norm
        fd = s["fd"]
        dfw = dfb = None
        if fd is not None:
            fmod = enc.fddts[i]
            any_grad = any(p.requires_grad for p in fmod.parameters())
            if any_grad:
                dfw = torch.zeros(4, d, dtype=torch.float32, device=dev)
                dfb = torch.zeros(4, d, dtype=torch.float32, device=dev)
        rows_in, stno = s["rows_in"], s["stno_in"]
        dgam = g.get(n1.weight) if g.want(n1.weight) else None
        dbet = g.get(n1.bias) if g.want(n1.bias) else None
        G = torch.empty(rows_in, d, dtype=torch.float32, device=dev)
        Gb = torch.empty(rows_in, d, dtype=torch.bfloat16, device=dev)
        # the Gb written below is dY of layer i - 1's fc2: reserve that layer's gradient bucket now so that the kernel can add
        # the bias gradient (column sums) while it writes the rows
        flat_next = g.reserve(group(i - 1)) if i > 0 else None
        gs = fc2_bias_grad(i - 1) if (i > 0 and tape.layers[i - 1]["B"] * T == rows_in and s["full"] is None) else None
        fc2_bias_done = gs is not None
        if s["scb"] is None:
            ops.layernorm_fddt_bwd(s["x_pre"], G, dy=dln1, g_in=G2, gamma=e["ln1_g"], delta1=s["d1_in"], delta2=s["d2_in"],
                                   T=T, stno=stno if fd is not None else None, fddt_w=fd[0] if fd is not None else None,
                                   fddt_b=fd[1] if fd is not None else None, g_out_bf16=Gb, dgamma=dgam, dbeta=dbet,
                                   dfddt_w=dfw, dfddt_b=dfb, g_colsum=gs)
        else:
            # LayerNorm 1 alone -> gradient of the stream after the speaker communication block; enrollment rows that
            # were dropped after this layer (encoder.py:210-213) carry no gradient from above
            Gs = torch.empty(rows_in, d, dtype=torch.float32, device=dev)
            if rows_in != rows:
                Gs[rows:].zero_()
            ops.layernorm_fddt_bwd(s["x_post"], Gs[:rows], dy=dln1, g_in=G2, gamma=e["ln1_g"], dgamma=dgam, dbeta=dbet)
            _scb_backward(enc, g, enc.ca_enrolls[i].cae, w["scb"][i], s["scb"], Gs, rows_in // (2 * T), T)
            ops.layernorm_fddt_bwd(s["x_pre"], G, g_in=Gs, delta1=s["d1_in"], delta2=s["d2_in"], T=T,
                                   stno=stno if fd is not None else None, fddt_w=fd[0] if fd is not None else None,
                                   fddt_b=fd[1] if fd is not None else None, g_out_bf16=Gb, dfddt_w=dfw, dfddt_b=dfb,
                                   g_colsum=gs)
        if dfw is not None:
            _scatter_fddt_grads(g, enc.fddts[i], dfw, dfb)
        if s["full"] is not None:  # G is the gradient of the TRANSFORMED stream: back through the four class transforms
            G, Gb = _fddt_full_backward(g, enc.fddts[i], G, stno, s["full"]["xb"], s["full"]["W4"], T)
        g.flush(flat)
        tape.layers[i] = None  # this layer's activations are dead: let the allocator reuse them
    if not n_layers:
        g.flush(flat)
    flat = g.reserve([enc.conv1.weight, enc.conv1.bias, enc.conv2.weight, enc.conv2.bias, enc.embed_positions.weight]
                     + (list(enc.initial_fddt.parameters()) if (cfg.use_fddt and cfg.use_pre_pos_fddt) else []))
    _stem_backward(enc, g, G, tape)
    _fold_conv_grads(g)
    g.flush(flat)


def _fddt_full_backward(g: _Grads, fmod, G: torch.Tensor, stno: torch.Tensor, xb: torch.Tensor, W4: torch.Tensor, T: int):
    """Backward of x' = sum_c mask_c * CustomLinear_c(x) (full-matrix FDDT, src/models/dicow/FDDT.py:52-62) computed as one
    stacked projection: G fp32 [rows, d] = dL/dx' -> (dL/dx fp32, its bf16 copy); class weight / bias gradients accumulated.
    A disabled class is an identity block of W4 and only contributes to dx."""
    rows, d = G.shape
    dY = torch.empty(rows, 4 * d, dtype=torch.bfloat16, device=G.device)
    ops.fddt_full_scatter(G, stno, dY, T=T)
    dx = torch.zeros(rows, d, dtype=torch.float32, device=G.device)
    ops.gemm(dY, W4, dx, epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_W_T, splits=1)
    for c, name in enumerate(_FDDT_ORDER):
        lin = getattr(fmod, name + "_linear", None)
        if lin is None:
            continue
        blk = slice(c * d, (c + 1) * d)
        _linear_backward(g, dY, xb, W4, lin.weight, lin.bias, need_dx=False, dy_cols=blk, w_rows=blk)
    return dx, ops.cast_bf16(dx)


def _scatter_fddt_grads(g: _Grads, fmod, dfw: torch.Tensor, dfb: torch.Tensor) -> None:
    for c, name in enumerate(_FDDT_ORDER):
        lin = getattr(fmod, name + "_linear", None)
        if lin is None:
            continue
        if isinstance(lin, torch.nn.Parameter):  # bias-only FDDT: the parameter is the class bias (FDDT.py:10)
            if g.want(lin):
                g.get(lin).add_(dfb[c])
            continue
        if g.want(lin.weight):
            g.get(lin.weight).add_(dfw[c])
        if lin.bias is not None and g.want(lin.bias):
            g.get(lin.bias).add_(dfb[c])


def _stem_backward(enc, g: _Grads, G0: torch.Tensor, tape: EncoderTape) -> None:
    """conv1 + GELU, conv2 + GELU, initial FDDT, + positions (encoder.py:167-179).  The forward fuses GELU / FDDT / pos
    into the conv2 epilogue; the backward recomputes conv2's pre-activation once."""
    cfg = enc.config
    stem_params = [enc.conv1.weight, enc.conv1.bias, enc.conv2.weight, enc.conv2.bias]
    has_f0 = cfg.use_fddt and cfg.use_pre_pos_fddt
    f0_params = list(enc.initial_fddt.parameters()) if has_f0 else []
    if not any(p.requires_grad for p in stem_params + f0_params + [enc.embed_positions.weight]):
        return
    w = enc.prepare()
    s = tape.stem
    B, T, F, d, C = s["B"], s["T"], s["F"], cfg.d_model, cfg.num_mel_bins
    dev = G0.device
    if g.want(enc.embed_positions.weight):  # reference quirk: unfreezing flips it to trainable
        g.get(enc.embed_positions.weight).add_(G0.view(B, T, d).sum(0))
    # recompute gelu(conv2) and its pre-activation
    g2 = torch.empty(B * T, d, dtype=torch.bfloat16, device=dev)
    pre2 = torch.empty(B * T, d, dtype=torch.bfloat16, device=dev)
    ops.gemm(s["a1"], w["conv2_w"], g2, epilogue=ops.EPI_GELU_SAVE_BF16, bias=w["conv2_b"], nb=B, Mb=T, K=3 * d, lda=2 * d,
             a_batch_stride=(F + 2) * d, ldo=d, out_batch_stride=T * d, aux=pre2)
    fw0, fb0 = w["fddt0"]
    if s.get("full"):
        dg2, _ = _fddt_full_backward(g, enc.initial_fddt, G0, s["stno0"], g2, w["fddt0_full"][0], T)
    else:
        dfw = torch.zeros(4, d, dtype=torch.float32, device=dev) if has_f0 else None
        dfb = torch.zeros(4, d, dtype=torch.float32, device=dev) if has_f0 else None
        dg2 = torch.empty(B * T, d, dtype=torch.float32, device=dev)
        zero_x = torch.zeros(B * T, d, dtype=torch.float32, device=dev)
        ops.layernorm_fddt_bwd(zero_x, dg2, g_in=G0, delta1=g2, T=T, stno=s["stno0"], fddt_w=fw0, fddt_b=fb0, dfddt_w=dfw,
                               dfddt_b=dfb)
        if has_f0:
            _scatter_fddt_grads(g, enc.initial_fddt, dfw, dfb)
    if not any(p.requires_grad for p in stem_params):
        return
    dpre2 = ops.dgelu_mul(dg2, pre2)  # bf16 [B*T, d]
    need_dx = enc.conv1.weight.requires_grad or enc.conv1.bias.requires_grad
    da1 = _conv_backward(g, dpre2.view(B, T, d), s["a1"], w["conv2_w"], enc.conv2.weight, enc.conv2.bias, B=B, T_in=F,
                         T_out=T, C_in=d, stride=2, need_dx=need_dx, dx_padded=True)
    if need_dx:
        # da1 / pre1 share the zero-padded [B, F + 2, d] layout of the forward buffer (padding rows carry zero gradient)
        dpre1 = ops.dgelu_mul(da1.view(B * (F + 2), d), s["pre1"].view(B * (F + 2), d))
        _conv_backward(g, dpre1.view(B, F + 2, d)[:, 1:F + 1], s["a0"], w["conv1_w"], enc.conv1.weight, enc.conv1.bias, B=B,
                       T_in=F, T_out=F, C_in=C, stride=1, need_dx=False, bias_src=dpre1)


# ----------------------------------------------------------------------------------------------------------------
# decoder (teacher forced)
# ----------------------------------------------------------------------------------------------------------------
class DecoderTape:
    def __init__(self):
        self.layers: List[dict] = []
        self.final: dict = {}


def _self_attn_strides(S: int, d: int) -> dict:
    return dict(q_row_stride=3 * d, q_batch_stride=S * 3 * d, kv_row_stride=3 * d, kv_batch_stride=S * 3 * d,
                o_row_stride=d, o_batch_stride=S * d)


def _cross_attn_strides(S: int, T: int, d: int) -> dict:
    return dict(q_row_stride=d, q_batch_stride=S * d, kv_row_stride=2 * d, kv_batch_stride=T * 2 * d, o_row_stride=d,
                o_batch_stride=S * d)


def decoder_forward_train(model, decoder_input_ids: torch.Tensor, enc_bf16: torch.Tensor, B: int, T: int, tape: DecoderTape):
    """HF WhisperDecoder.forward without cache (HF:modeling_whisper.py:691-796) keeping the backward's inputs; the same
    kernels as DiCoW.decode_teacher_forced, residual GEMMs writing a new stream buffer instead of updating in place.
    enc_bf16 [B*T, d].  Returns (hidden fp32 [B, S, d], hidden bf16 [B*S, d])."""
    cfg = model.config
    w = model.prepare_decoder()
    dev = enc_bf16.device
    S = decoder_input_ids.shape[1]
    d, H, ffn = cfg.d_model, cfg.decoder_attention_heads, cfg.decoder_ffn_dim
    if S > cfg.max_target_positions:
        raise ValueError(f"decoder sequence length {S} exceeds max_target_positions {cfg.max_target_positions}")
    ids = decoder_input_ids.to(device=dev, dtype=torch.int64).contiguous()
    rows = B * S
    x = torch.empty(rows, d, dtype=torch.float32, device=dev)
    ops.embed_tokens(ids, w["tok"], w["pos"], x.view(B, S, d), S=S, past=0)

    def bf(*shape):
        return torch.empty(*shape, dtype=torch.bfloat16, device=dev)

    for e in w["layers"]:
        s, c = e["self"], e["cross"]
        x0 = x
        ln1 = bf(rows, d)
        ops.fddt_layernorm(x0, gamma=e["ln1_g"], beta=e["ln1_b"], ln_out_bf16=ln1)
        qkv = bf(rows, 3 * d)
        ops.gemm(ln1, s["wqkv"], qkv, epilogue=ops.EPI_BIAS_BF16, bias=s["bqkv"])
        ctx_s, lse_s = bf(rows, d), torch.empty(B, H, S, dtype=torch.float32, device=dev)
        ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], ctx_s, B=B, H=H, Tq=S, Tk=S, causal=True, lse=lse_s,
                      **_self_attn_strides(S, d))
        x1 = torch.empty(rows, d, dtype=torch.float32, device=dev)
        ops.gemm(ctx_s, s["wo"], x1, epilogue=ops.EPI_RESIDUAL_F32, bias=s["bo"], resid=x0)
        ln2 = bf(rows, d)
        ops.fddt_layernorm(x1, gamma=e["ln2_g"], beta=e["ln2_b"], ln_out_bf16=ln2)
        q = bf(rows, d)
        ops.gemm(ln2, c["wq"], q, epilogue=ops.EPI_BIAS_BF16, bias=c["bq"])
        kv = bf(B * T, 2 * d)
        ops.gemm(enc_bf16, c["wkv"], kv, epilogue=ops.EPI_BIAS_BF16, bias=c["bkv"])
        ctx_c, lse_c = bf(rows, d), torch.empty(B, H, S, dtype=torch.float32, device=dev)
        ops.attention(q, kv, kv[:, d:], ctx_c, B=B, H=H, Tq=S, Tk=T, lse=lse_c, **_cross_attn_strides(S, T, d))
        x2 = torch.empty(rows, d, dtype=torch.float32, device=dev)
        ops.gemm(ctx_c, c["wo"], x2, epilogue=ops.EPI_RESIDUAL_F32, bias=c["bo"], resid=x1)
        ln3 = bf(rows, d)
        ops.fddt_layernorm(x2, gamma=e["ln3_g"], beta=e["ln3_b"], ln_out_bf16=ln3)
        hdn, pre = bf(rows, ffn), bf(rows, ffn)
        ops.gemm(ln3, e["w1"], hdn, epilogue=ops.EPI_GELU_SAVE_BF16, bias=e["b1"], aux=pre)
        x3 = torch.empty(rows, d, dtype=torch.float32, device=dev)
        ops.gemm(hdn, e["w2"], x3, epilogue=ops.EPI_RESIDUAL_F32, bias=e["b2"], resid=x2)
        tape.layers.append({"x0": x0, "ln1": ln1, "qkv": qkv, "ctx_s": ctx_s, "lse_s": lse_s, "x1": x1, "ln2": ln2, "q": q,
                            "kv": kv, "ctx_c": ctx_c, "lse_c": lse_c, "x2": x2, "ln3": ln3, "hdn": hdn, "pre": pre})
        x = x3
    hid = torch.empty(B, S, d, dtype=torch.float32, device=dev)
    hid_bf16 = bf(rows, d)
    ops.fddt_layernorm(x, gamma=w["lnf_g"], beta=w["lnf_b"], ln_out_f32=hid, ln_out_bf16=hid_bf16)
    tape.final = {"x": x, "ids": ids, "enc": enc_bf16, "B": B, "S": S, "T": T, "hid_bf16": hid_bf16}
    return hid, hid_bf16


def _ln_bwd(g: _Grads, ln_mod, x, gamma, dy, g_in, rows, d, dev):
    G = torch.empty(rows, d, dtype=torch.float32, device=dev)
    Gb = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
    ops.layernorm_fddt_bwd(x, G, dy=dy, g_in=g_in, gamma=gamma, g_out_bf16=Gb,
                           dgamma=g.get(ln_mod.weight) if g.want(ln_mod.weight) else None,
                           dbeta=g.get(ln_mod.bias) if g.want(ln_mod.bias) else None)
    return G, Gb


def decoder_backward(model, g: _Grads, dlogits: torch.Tensor, tape: DecoderTape, d_enc_accum: Optional[torch.Tensor]):
    """dlogits bf16 [B*S, ldd] (ldd = ceil8(vocab), padding columns zero): gradients of the decoder parameters that
    require grad, and dL/d(encoder states) accumulated into the fp32 ``d_enc_accum`` [B*T, d] (skipped when None)."""
    cfg = model.config
    w = model.prepare_decoder()
    dec = model.decoder
    f = tape.final
    B, S, T, d, H = f["B"], f["S"], f["T"], cfg.d_model, cfg.decoder_attention_heads
    rows = B * S
    dev = dlogits.device
    V, ldd = w["proj"].shape[0], dlogits.shape[1]
    # proj_out (tied to embed_tokens, no bias): d_hid = dlogits W ; dW += dlogits^T hid
    dhid = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
    ops.gemm(dlogits, w["proj"], dhid, epilogue=ops.EPI_BIAS_BF16, flags=ops.GEMM_W_T, K=V, lda=ldd, Mb=rows)
    emb = dec.embed_tokens.weight
    flat_top = g.reserve([emb, dec.embed_positions.weight, dec.layer_norm.weight, dec.layer_norm.bias])
    if g.want(emb):
        ops.gemm(dlogits, f["hid_bf16"], g.get_padded(emb), epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_A_T | ops.GEMM_W_T,
                 lda=ldd, Mb=ldd, K=rows, N=d)
    G, Gb = _ln_bwd(g, dec.layer_norm, f["x"], w["lnf_g"], dhid, None, rows, d, dev)
    sc = 64 ** -0.5
    for i in range(len(w["layers"]) - 1, -1, -1):
        e, t, lyr = w["layers"][i], tape.layers[i], dec.layers[i]
        s, c = e["self"], e["cross"]
        flat = g.reserve(list(lyr.parameters()))
        # MLP
        dpre = _linear_backward(g, Gb, t["hdn"], e["w2"], lyr.fc2.weight, lyr.fc2.bias, dx_epilogue=ops.EPI_DGELU_BF16,
                                aux=t["pre"])
        dln3 = _linear_backward(g, dpre, t["ln3"], e["w1"], lyr.fc1.weight, lyr.fc1.bias)
        G2, G2b = _ln_bwd(g, lyr.final_layer_norm, t["x2"], e["ln3_g"], dln3, G, rows, d, dev)
        # cross attention
        ca = lyr.encoder_attn
        dctx = _linear_backward(g, G2b, t["ctx_c"], c["wo"], ca.out_proj.weight, ca.out_proj.bias)
        dq = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        dkv = torch.empty(B * T, 2 * d, dtype=torch.bfloat16, device=dev)
        kv = t["kv"]
        ops.attention_bwd(t["q"], kv, kv[:, d:], t["ctx_c"], dctx, t["lse_c"], dq, dkv, dkv[:, d:], B=B, H=H, Tq=S, Tk=T,
                          dq_row_stride=d, dq_batch_stride=S * d, dkv_row_stride=2 * d, dkv_batch_stride=T * 2 * d,
                          **_cross_attn_strides(S, T, d))
        dln2 = _linear_backward(g, dq, t["ln2"], c["wq"], ca.q_proj.weight, ca.q_proj.bias, scale=sc)
        if d_enc_accum is not None:
            ops.gemm(dkv, c["wkv"], d_enc_accum, epilogue=ops.EPI_ACCUM_F32, flags=ops.GEMM_W_T, splits=1)
        _linear_backward(g, dkv, f["enc"], c["wkv"], ca.k_proj.weight, None, need_dx=False, dy_cols=slice(0, d),
                         w_rows=slice(0, d))
        _linear_backward(g, dkv, f["enc"], c["wkv"], ca.v_proj.weight, ca.v_proj.bias, need_dx=False, dy_cols=slice(d, 2 * d),
                         w_rows=slice(d, 2 * d))
        G1, G1b = _ln_bwd(g, lyr.encoder_attn_layer_norm, t["x1"], e["ln2_g"], dln2, G2, rows, d, dev)
        # causal self attention
        sa = lyr.self_attn
        dctx = _linear_backward(g, G1b, t["ctx_s"], s["wo"], sa.out_proj.weight, sa.out_proj.bias)
        dqkv = torch.empty(rows, 3 * d, dtype=torch.bfloat16, device=dev)
        qkv = t["qkv"]
        ops.attention_bwd(qkv, qkv[:, d:], qkv[:, 2 * d:], t["ctx_s"], dctx, t["lse_s"], dqkv, dqkv[:, d:], dqkv[:, 2 * d:],
                          B=B, H=H, Tq=S, Tk=S, causal=True, dq_row_stride=3 * d, dq_batch_stride=S * 3 * d,
                          dkv_row_stride=3 * d, dkv_batch_stride=S * 3 * d, **_self_attn_strides(S, d))
        dln1 = _attention_params_backward(g, sa, s, dqkv, t["ln1"], d, True)
        G, Gb = _ln_bwd(g, lyr.self_attn_layer_norm, t["x0"], e["ln1_g"], dln1, G1, rows, d, dev)
        g.flush(flat)
    pos = dec.embed_positions.weight
    if g.want(emb) or g.want(pos):
        ops.embedding_bwd(G, f["ids"].view(-1), S=S, d_tok=g.get(emb) if g.want(emb) else None,
                          d_pos=g.get(pos) if g.want(pos) else None)
    g.flush(flat_top)


# ----------------------------------------------------------------------------------------------------------------
# autograd bridges
# ----------------------------------------------------------------------------------------------------------------
_HEAD_PREFIXES = ("additional_self_attention_layer", "subsample_conv", "lm_head", "additional_layer")


def _body_trainable(enc) -> bool:
    return any(p.requires_grad for n, p in enc.named_parameters() if not n.startswith(_HEAD_PREFIXES))


_STOCHASTIC = ("dropout", "attention_dropout", "activation_dropout", "final_dropout", "encoder_layerdrop", "decoder_layerdrop")


def refuse_stochastic_regularisation(config) -> None:
    """The training step applies no dropout / LayerDrop / in-model SpecAugment: every Whisper checkpoint and the reference's
    DiCoWConfig default have them at zero (config.py:14; apply_spec_augment False).  A config that asks for them must not be
    trained silently without them."""
    asked = [f"{k}={getattr(config, k)}" for k in _STOCHASTIC if float(getattr(config, k, 0.0) or 0.0) > 0.0]
    if getattr(config, "apply_spec_augment", False):
        asked.append("apply_spec_augment=True")
    if asked:
        raise NotImplementedError("the B200 training step does not implement: " + ", ".join(asked))


def trainable(module: torch.nn.Module) -> bool:
    """route forward() to the training path: autograd is recording and something can receive a gradient"""
    on = torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())
    if on and module.training and getattr(module, "config", None) is not None:
        refuse_stochastic_regularisation(module.config)
    return on


def _enrollments(enr_features, enr_stno) -> Optional[dict]:
    return None if enr_features is None else {"input_features": enr_features, "stno_mask": enr_stno}


def _encoder_hidden(enc, input_features, stno_mask, tape: EncoderTape, body: bool, enrollments: Optional[dict] = None):
    """(hidden fp32 [B, T, d], hidden bf16 [B*T, d], B, T); activations are saved only when the body trains."""
    if body:
        hidden, hidden_bf16 = encoder_forward_train(enc, input_features, stno_mask, tape, enrollments)
        return hidden, hidden_bf16, tape.final["B"], tape.final["T"]
    out = enc(input_features, stno_mask=stno_mask, enrollments=enrollments)  # frozen body: inference path, nothing saved
    hidden = out.last_hidden_state
    B, T = hidden.shape[0], hidden.shape[1]
    return hidden, ops.cast_bf16(hidden).view(B * T, -1), B, T


class EncoderLogitsFn(torch.autograd.Function):
    """DiCoWEncoder.forward(return_logits=True) with a grad_fn (CTC pre-training: src/pretrain_encoder.py:42-51 freezes
    everything but the CTC head, src/utils/trainers.py:76-103 then calls model.get_loss on these logits)."""

    @staticmethod
    def forward(ctx, enc, input_features, stno_mask, enr_features, enr_stno, *params):
        if gradient_exchange is not None:
            gradient_exchange.ensure_synced(enc)
        tape = EncoderTape()
        body = _body_trainable(enc)
        ctx.epoch = ops.prepare_epoch = ops.new_prepare_epoch()
        try:
            with torch.no_grad():
                hidden, hidden_bf16, B, T = _encoder_hidden(enc, input_features, stno_mask, tape, body,
                                                            _enrollments(enr_features, enr_stno))
                logits = ctc_head_forward_train(enc, hidden_bf16, B, T, tape)
        finally:
            ops.prepare_epoch = 0
        ctx.enc, ctx.tape, ctx.body, ctx.params = enc, tape, body, params
        neck = tape.ctc["neck"].float()  # what the reference returns as hidden_states (encoder.py:233-240)
        ctx.mark_non_differentiable(hidden, neck)
        return logits, hidden, neck

    @staticmethod
    def backward(ctx, grad_logits, _grad_hidden, _grad_neck):
        enc, tape = ctx.enc, ctx.tape
        g = _Grads(gradient_exchange)
        ops.prepare_epoch = ctx.epoch  # the prepared weights of this step's forward
        try:
            with torch.no_grad():
                V1 = grad_logits.shape[-1]
                dl = ops.cast_bf16_padded(grad_logits.reshape(-1, V1).float(), _ceil8(V1))
                dh = ctc_head_backward(enc, g, dl.view(grad_logits.shape[0], grad_logits.shape[1], -1), tape,
                                       need_dhidden=ctx.body)
                if ctx.body:
                    encoder_backward(enc, g, dh, tape)
        finally:
            ops.prepare_epoch = 0
        ctx.tape = None
        return (None,) * 5 + g.finish(ctx.params)


class CtcLossFn(torch.autograd.Function):
    """DiCoWEncoder.get_loss on logits that carry a grad_fn (src/models/dicow/encoder.py:108-135)."""

    @staticmethod
    def forward(ctx, logits, labels, reduction):
        with torch.no_grad():
            lg = ops._ctc_rows(logits.float())
            loss, ws = ops.ctc_loss_with_lse(lg, labels, reduction)
        ctx.save_for_backward(lg, labels, ws)
        ctx.reduction = reduction
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lg, labels, ws = ctx.saved_tensors
        with torch.no_grad():
            gl = grad_loss.reshape(1).float().contiguous()
            dl = ops.ctc_loss_bwd(lg, labels, ws, ctx.reduction, scale_dev=gl, out_f32=True)
        return dl, None, None


class DiCoWTrainStepFn(torch.autograd.Function):
    """loss of DiCoWForConditionalGeneration.forward (src/models/dicow/modeling_dicow.py:248-354):
    (1 - w) * soft-label CE of the teacher-forced decoder + w * CTC of the encoder head, with one hand-scheduled backward."""

    @staticmethod
    def forward(ctx, model, input_features, stno_mask, decoder_input_ids, labels, upp_labels, enc_labels, enr_features,
                enr_stno, *params):
        cfg = model.config
        if gradient_exchange is not None:
            gradient_exchange.ensure_synced(model)
        enc = model.model.get_encoder()
        etape, dtape = EncoderTape(), DecoderTape()
        # the encoder states receive a gradient from the decoder's cross-attention whenever the body trains
        body = _body_trainable(enc)
        dec_trainable = any(p.requires_grad for p in model.model.decoder.parameters())
        ctx.epoch = ops.prepare_epoch = ops.new_prepare_epoch()
        try:
            return DiCoWTrainStepFn._forward(ctx, model, enc, etape, dtape, body, dec_trainable, input_features, stno_mask,
                                             decoder_input_ids, labels, upp_labels, enc_labels, enr_features, enr_stno, params)
        finally:
            ops.prepare_epoch = 0

    @staticmethod
    def _forward(ctx, model, enc, etape, dtape, body, dec_trainable, input_features, stno_mask, decoder_input_ids, labels,
                 upp_labels, enc_labels, enr_features, enr_stno, params):
        cfg = model.config
        with torch.no_grad():
            hidden, hidden_bf16, B, T = _encoder_hidden(enc, input_features, stno_mask, etape, body,
                                                        _enrollments(enr_features, enr_stno))
            dev = hidden.device
            if body or dec_trainable:
                hid, hid_bf16 = decoder_forward_train(model.model, decoder_input_ids, hidden_bf16, B, T, dtape)
            else:
                hid, hid_bf16 = model.model.decode_teacher_forced(decoder_input_ids, hidden_bf16.view(B, T, -1))
            S = decoder_input_ids.shape[1]
            w = model.model.prepare_decoder()
            V = cfg.vocab_size
            logits = torch.empty(B, S, V, dtype=torch.float32, device=dev)
            ops.gemm(hid_bf16, w["proj"], logits.view(B * S, V), epilogue=ops.EPI_BIAS_F32)
            labels = labels.to(dev).contiguous()
            upp = upp_labels.to(dev).contiguous() if upp_labels is not None else None
            slc = model.soft_label_creator
            ce = dict(ts_begin=slc.ts_begin if slc is not None else 0,
                      smoothing=slc.smoothing_on(dev) if slc is not None else None,
                      soft_mode=slc is not None)
            dec_loss = ops.softlabel_ce(logits.view(B * S, V), labels, upp, **ce)
            wctc = float(cfg.ctc_weight)
            ctc_ws = enc_logits = None
            if wctc > 0.0:
                enc_logits = ctc_head_forward_train(enc, hidden_bf16, B, T, etape)
                if callable(enc_labels):  # deferred host-side bookkeeping (DiCoWForConditionalGeneration.forward)
                    enc_labels = enc_labels()
                enc_labels = enc_labels.to(dev).contiguous()
                ctc, ctc_ws = ops.ctc_loss_with_lse(enc_logits, enc_labels, cfg.ctc_loss_reduction)
                loss = (1 - wctc) * dec_loss + wctc * ctc
            else:
                loss = dec_loss
            # normaliser of the decoder loss: non-padding tokens (soft labels) or all rows (hard-label fallback)
            n_norm = (labels != -100).sum().float() if slc is not None else torch.tensor(float(B * S), device=dev)
        ctx.model, ctx.etape, ctx.dtape, ctx.params = model, etape, dtape, params
        ctx.body, ctx.dec_trainable, ctx.ce, ctx.wctc = body, dec_trainable, ce, wctc
        ctx.saved = (logits, labels, upp, enc_logits, enc_labels if wctc > 0.0 else None, ctc_ws, n_norm)
        ctx.mark_non_differentiable(logits, hidden)
        return loss, logits, hidden

    @staticmethod
    def backward(ctx, grad_loss, _gl, _gh):
        model = ctx.model
        cfg = model.config
        enc = model.model.get_encoder()
        logits, labels, upp, enc_logits, enc_labels, ctc_ws, n_norm = ctx.saved
        g = _Grads(gradient_exchange)
        B, S, V = logits.shape
        ops.prepare_epoch = ctx.epoch  # the prepared weights of this step's forward
        try:
            DiCoWTrainStepFn._backward(ctx, model, cfg, enc, g, grad_loss, logits, labels, upp, enc_logits, enc_labels, ctc_ws,
                                       n_norm, B, S, V)
        finally:
            ops.prepare_epoch = 0
        ctx.etape = ctx.dtape = ctx.saved = None
        return (None,) * 9 + g.finish(ctx.params)

    @staticmethod
    def _backward(ctx, model, cfg, enc, g, grad_loss, logits, labels, upp, enc_logits, enc_labels, ctc_ws, n_norm, B, S, V):
        with torch.no_grad():
            gl = grad_loss.reshape(1).float()
            d_enc = None
            if ctx.body or ctx.dec_trainable:
                s_dec = (gl * (1.0 - ctx.wctc) / n_norm.clamp(min=1.0)).contiguous()
                dlogits = ops.softlabel_ce_bwd(logits.view(B * S, V), labels, upp, scale=1.0, scale_dev=s_dec, **ctx.ce)
                if ctx.body:
                    T = ctx.dtape.final["T"]
                    d_enc = torch.zeros(B * T, cfg.d_model, dtype=torch.float32, device=logits.device)
                decoder_backward(model.model, g, dlogits, ctx.dtape, d_enc)
                del dlogits
            if ctx.wctc > 0.0:
                dl = ops.ctc_loss_bwd(enc_logits, enc_labels, ctc_ws, cfg.ctc_loss_reduction, loss_scale=ctx.wctc,
                                      scale_dev=gl.contiguous())
                ctc_head_backward(enc, g, dl, ctx.etape, need_dhidden=ctx.body, dhidden_accum=d_enc)
                del dl
            if ctx.body:
                encoder_backward(enc, g, ops.cast_bf16(d_enc), ctx.etape)
