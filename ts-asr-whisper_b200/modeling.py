"""Host-side mirror of the reference's model classes for the hot path (encoder side).

Same class names, constructor arguments, ``forward`` signatures and state_dict parameter names as
  src/models/dicow/encoder.py  (DiCoWEncoder)           src/models/dicow/FDDT.py   (FDDT)
  src/models/dicow/layers.py   (CustomDiagonalLinear, Gate, CrossAttentionEnrollBlock, SpeakerCommunicationBlock)
so checkpoints, name-keyword freezing (src/models/containers.py:80-97) and HF ``generate()``'s signature
inspection keep working -- but the modules are parameter containers only: every FLOP of ``forward`` is a call into
libdicow_b200.so (ops.py).  There is no eager / CPU fallback: without the library or an sm_100 GPU, forward raises.

Inference (no-grad) path; the residual stream is fp32, GEMM / attention operands bf16 with fp32 accumulation, which
is the reference's own bf16-autocast numerics (SURVEY.md Appendix A "Training numerics", section 7 "Mixed-precision").
"""
from __future__ import annotations

import os
import weakref

import math
from typing import Dict, Optional

import torch
from torch import nn
from dataclasses import dataclass

from transformers.modeling_outputs import BaseModelOutput, CausalLMOutput


@dataclass
class DiCoWCTCOutput(CausalLMOutput):
    """``DiCoWEncoder.forward(return_logits=True)``: the reference's CausalLMOutput (encoder.py:233-240) -- ``hidden_states``
    is the CTC neck output it feeds to lm_head ([B, T/4, d] with pre_ctc_sub_sample) -- plus the final-LayerNorm output the
    reference discards on this path (kept because the joint decoding / training callers here reuse it)."""
    encoder_last_hidden_state: Optional[torch.FloatTensor] = None

from . import ops
from .configuration import DiCoWConfig

_FDDT_ORDER = ("silence", "target", "non_target", "overlap")  # STNO channel order (FDDT.py:56-62)


# ----------------------------------------------------------------------------------------------------------------
# parameter containers (names == reference state_dict)
# ----------------------------------------------------------------------------------------------------------------
class CustomDiagonalLinear(nn.Module):
    """weight/bias [d]; inits follow src/models/dicow/layers.py:49-71."""

    def __init__(self, d_model: int, bias: bool = True, init_eye_val: float = 0.0, fddt_init: Optional[str] = None):
        super().__init__()
        self.init_eye_val = init_eye_val
        self.fddt_init = fddt_init
        self.weight = nn.Parameter(torch.empty(d_model))
        self.bias = nn.Parameter(torch.zeros(d_model)) if bias else None
        self.reset_parameters()

    def reset_parameters(self, weight: bool = True, bias: bool = True) -> None:
        """``weight`` / ``bias`` = False leaves that tensor alone (it came from a checkpoint)"""
        with torch.no_grad():
            if weight:
                bound = math.sqrt(3.0 / self.weight.numel())
                self.weight.uniform_(-bound, bound)
                if self.fddt_init == "non-disturbing":
                    self.weight.fill_(1.0)
                elif self.fddt_init == "suppressive":
                    self.weight.fill_(self.init_eye_val)
            if bias and self.bias is not None:
                self.bias.zero_()


class CustomLinear(nn.Linear):
    """nn.Linear with the FDDT inits of src/models/dicow/layers.py:7-47 (full d x d class transform)."""

    def __init__(self, *args, init_eye_val: float = 0.0, fddt_init: Optional[str] = None, **kwargs):
        self.init_eye_val, self.fddt_init = init_eye_val, fddt_init
        super().__init__(*args, **kwargs)

    def reset_parameters(self, weight: bool = True, bias: bool = True) -> None:
        with torch.no_grad():
            if weight:
                nn.init.xavier_uniform_(self.weight)
                scale = {"non-disturbing": 1.0, "suppressive": self.init_eye_val}.get(self.fddt_init)
                if scale is not None:
                    n = min(self.weight.shape)
                    self.weight.zero_()
                    self.weight[:n, :n] = scale * torch.eye(n, device=self.weight.device)
            if bias and self.bias is not None:
                self.bias.zero_()


class FDDT(nn.Module):
    """Frame-level Diarization-Dependent Transformation parameters (src/models/dicow/FDDT.py:6-40): per STNO class a
    diagonal transform (default), a full d x d one (``is_diagonal=False``) or a bias vector (``bias_only``)."""

    def __init__(self, d_model: int, non_target_rate: float = 0.01, fddt_init: Optional[str] = None,
                 is_diagonal: bool = True, bias_only: bool = False, use_silence: bool = True, use_target: bool = True,
                 use_overlap: bool = True, use_non_target: bool = True):
        super().__init__()
        self.d_model, self.is_diagonal, self.bias_only = d_model, is_diagonal, bias_only

        def make(eye_val: float):
            if bias_only:
                return nn.Parameter(torch.zeros(d_model))
            if is_diagonal:
                return CustomDiagonalLinear(d_model, True, eye_val, fddt_init)
            return CustomLinear(d_model, d_model, bias=True, fddt_init=fddt_init, init_eye_val=eye_val)

        if use_target:
            self.target_linear = make(1.0)
        if use_non_target:
            self.non_target_linear = make(non_target_rate)
        if use_overlap:
            self.overlap_linear = make(1.0)
        if use_silence:
            self.silence_linear = make(non_target_rate)

    def tables(self):
        """([4, d] weights or None for the bias-only variant, [4, d] biases) in STNO order; a disabled class is the
        identity (FDDT.py:54-62).  Diagonal / bias-only variants."""
        ref = next(self.parameters())
        ws, bs = [], []
        for c in _FDDT_ORDER:
            lin = getattr(self, c + "_linear", None)
            if self.bias_only:
                bs.append(lin if lin is not None else torch.zeros(self.d_model, device=ref.device))
                continue
            ws.append(lin.weight if lin is not None else torch.ones(self.d_model, device=ref.device))
            bs.append(lin.bias if lin is not None else torch.zeros(self.d_model, device=ref.device))
        w = None if self.bias_only else torch.stack(ws).detach().float().contiguous()
        return w, torch.stack(bs).detach().float().contiguous()

    def full_tables(self):
        """full-matrix variant: the four class transforms stacked into ONE projection (bf16 [4 d, d], fp32 bias [4 d]) in
        STNO order; a disabled class is the identity"""
        ref = next(self.parameters())
        d = self.d_model
        ws, bs = [], []
        for c in _FDDT_ORDER:
            lin = getattr(self, c + "_linear", None)
            ws.append(lin.weight.detach().float() if lin is not None else torch.eye(d, device=ref.device))
            bs.append(lin.bias.detach().float() if lin is not None else torch.zeros(d, device=ref.device))
        return ops.cast_bf16(torch.cat(ws, 0).contiguous()), torch.cat(bs).contiguous()


class Gate(nn.Module):
    def __init__(self, items: int, init_val: float = 0.0):
        super().__init__()
        self.init_val = init_val
        self.gate = nn.Parameter(torch.full((items,), init_val))

    def reset_parameters(self) -> None:
        with torch.no_grad():
            self.gate.fill_(self.init_val)


class AttentionParams(nn.Module):
    """q/k/v/out projections named like HF WhisperAttention (k_proj has no bias, HF:modeling_whisper.py:281-284)."""

    def __init__(self, d: int):
        super().__init__()
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d, bias=True)
        self.q_proj = nn.Linear(d, d, bias=True)
        self.out_proj = nn.Linear(d, d, bias=True)


class EncoderLayerParams(nn.Module):
    def __init__(self, d: int, ffn: int):
        super().__init__()
        self.self_attn = AttentionParams(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)


class CrossAttentionEnrollBlock(nn.Module):
    """Parameters of src/models/dicow/layers.py:113-143 (ffn is a Sequential so the names are ffn.0 / ffn.3)."""

    def __init__(self, config: DiCoWConfig):
        super().__init__()
        d, ffn = config.d_model, config.encoder_ffn_dim
        self.cross_attn = AttentionParams(d)
        self.cross_gate = Gate(1, init_val=0.0)
        self.ffn = nn.Sequential(nn.Linear(2 * d, ffn), nn.Identity(), nn.Identity(), nn.Linear(ffn, d), nn.Identity())
        self.ffn[0]._dicow_custom_init = self.ffn[3]._dicow_custom_init = True  # keep through HF post_init()
        self.reset_parameters()

    def reset_parameters(self) -> None:
        """layers.py:95-110: the update network starts as "copy the query stream through" (also re-run by
        DiCoWForConditionalGeneration._init_weights when from_pretrained() adds the block to a checkpoint without it)"""
        d = self.ffn[3].weight.shape[0]
        if self.ffn[0].weight.device.type == "meta":
            return
        with torch.no_grad():
            nn.init.xavier_uniform_(self.ffn[0].weight, gain=1e-1)
            self.ffn[0].weight[:d, :d] += torch.eye(d, device=self.ffn[0].weight.device)
            self.ffn[0].bias.zero_()
            nn.init.xavier_uniform_(self.ffn[3].weight, gain=1e-1)
            self.ffn[3].weight[:, :d] += torch.eye(d, device=self.ffn[3].weight.device)
            self.ffn[3].bias.zero_()


class SpeakerCommunicationBlock(nn.Module):
    def __init__(self, config: DiCoWConfig):
        super().__init__()
        self.streams = 2
        self.cae = CrossAttentionEnrollBlock(config)


# ----------------------------------------------------------------------------------------------------------------
# prepared (bf16 / fused) weights
# ----------------------------------------------------------------------------------------------------------------
# refresh the prepared weights of a training run through a captured CUDA graph (DiCoWEncoder.prepare); DICOW_PREPARE_GRAPH=0
# keeps the eager rebuild
prepare_graphs = os.environ.get("DICOW_PREPARE_GRAPH", "1") != "0"
_PREPARE_GRAPHS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()  # encoder module -> its captured refresh (kept out
#                                                                             of the module: CUDA graphs cannot be deep-copied)


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return ops.cast_bf16(t.detach().float())


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().float().contiguous()


def _prep_attention(att: AttentionParams, fuse_qkv: bool = True) -> Dict[str, torch.Tensor]:
    """bf16 projection weights; hd^-0.5 = 0.125 (exact in bf16 for head_dim 64) is folded into Wq / bq, which is the
    reference's ``q_proj(x) * scaling`` (HF:modeling_whisper.py:310) bit for bit in exact arithmetic."""
    d = att.q_proj.weight.shape[0]
    sc = 64 ** -0.5
    wq, bq = att.q_proj.weight.detach().float() * sc, att.q_proj.bias.detach().float() * sc
    wk = att.k_proj.weight.detach().float()
    wv, bv = att.v_proj.weight.detach().float(), att.v_proj.bias.detach().float()
    out = {"wo": _bf16(att.out_proj.weight), "bo": _f32(att.out_proj.bias)}
    if fuse_qkv:
        out["wqkv"] = _bf16(torch.cat([wq, wk, wv], 0))
        out["bqkv"] = torch.cat([bq, torch.zeros(d, device=bq.device), bv]).contiguous()
    else:
        out["wq"], out["bq"] = _bf16(wq), bq.contiguous()
        out["wkv"] = _bf16(torch.cat([wk, wv], 0))
        out["bkv"] = torch.cat([torch.zeros(d, device=bq.device), bv]).contiguous()
    return out


def _conv_weight(w: torch.Tensor) -> torch.Tensor:
    """Conv1d weight [Cout, Cin, 3] -> GEMM weight [Cout, 3*Cin] with k-major taps (matches the overlapping-row view
    of the zero-padded channels-last activation buffer)."""
    return _bf16(w.detach().float().permute(0, 2, 1).reshape(w.shape[0], -1))


# ----------------------------------------------------------------------------------------------------------------
# DiCoWEncoder
# ----------------------------------------------------------------------------------------------------------------
class DiCoWEncoder(nn.Module):
    """B200 implementation of src/models/dicow/encoder.py:10-246 (same attributes, forward signature, outputs)."""

    config_class = DiCoWConfig
    main_input_name = "input_features"

    def __init__(self, config: DiCoWConfig):
        super().__init__()
        config.check_supported()
        self.config = config
        d = config.d_model
        self.ctc_weight = config.ctc_weight
        self.conv1 = nn.Conv1d(config.num_mel_bins, d, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(d, d, kernel_size=3, stride=2, padding=1)
        self.embed_positions = nn.Embedding(config.max_source_positions, d)
        self.embed_positions.requires_grad_(False)
        self.layers = nn.ModuleList([EncoderLayerParams(d, config.encoder_ffn_dim)
                                     for _ in range(config.encoder_layers)])
        self.layer_norm = nn.LayerNorm(d)
        if config.additional_layer and self.ctc_weight > 0.0:  # encoder.py:16-17
            self.additional_layer = EncoderLayerParams(d, config.encoder_ffn_dim)
        if config.additional_self_attention_layer and self.ctc_weight > 0.0:
            self.additional_self_attention_layer = AttentionParams(d)
        if config.pre_ctc_sub_sample and self.ctc_weight > 0.0:
            self.subsample_conv1 = nn.Conv1d(d, d, kernel_size=3, stride=2, padding=1, bias=False)
            self.subsample_conv2 = nn.Conv1d(d, d, kernel_size=3, stride=2, padding=1, bias=False)
        if self.ctc_weight > 0.0:
            self.lm_head = nn.Linear(d, config.vocab_size + 1, bias=False)
        if config.use_fddt:
            n = config.apply_fddt_to_n_layers if config.apply_fddt_to_n_layers != -1 else len(self.layers)
            kw = dict(fddt_init=config.fddt_init, is_diagonal=config.fddt_is_diagonal, bias_only=config.fddt_bias_only,
                      use_silence=config.fddt_use_silence,
                      use_target=config.fddt_use_target, use_overlap=config.fddt_use_overlap,
                      use_non_target=config.fddt_use_non_target)
            self.fddts = nn.ModuleList([FDDT(d, non_target_rate=1.0, **kw) for _ in range(n)])
            if config.use_pre_pos_fddt:
                self.initial_fddt = FDDT(d, non_target_rate=config.non_target_fddt_value, **kw)
        if config.use_enrollments and config.scb_layers is not None:
            self.ca_enrolls = nn.ModuleList([SpeakerCommunicationBlock(config) for _ in range(config.scb_layers)])
        self.first_task_token = config.vocab_size - 30 * 50 - 1 - 6
        self._prepared: Optional[dict] = None
        self._prepared_key = None
        self.attention_variant = 0

    # ---- reference API surface -----------------------------------------------------------------------------
    def get_max_len(self) -> int:  # encoder.py:137-138
        return self.config.max_source_positions * self.conv1.stride[0] * self.conv2.stride[0]

    def get_output_embeddings(self):
        return None

    def get_loss(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """CTC loss as encoder.py:108-135 (fp32 log-softmax, blank = last class, every frame valid, zero_infinity) in
        the fused kernel (ops.ctc_loss); the label filtering is index bookkeeping on the int64 labels.  Logits that
        carry a grad_fn (training forward with return_logits=True) get the CTC backward kernels through autograd."""
        if labels.max() >= self.config.vocab_size:
            raise ValueError(f"Label values must be <= vocab_size: {self.config.vocab_size}")
        labels = self.ctc_label_filter(labels).to(logits.device).contiguous()
        if torch.is_grad_enabled() and logits.requires_grad:
            from .training import CtcLossFn
            return CtcLossFn.apply(logits, labels, self.config.ctc_loss_reduction)
        return ops.ctc_loss(logits.float(), labels, reduction=self.config.ctc_loss_reduction)

    def ctc_label_filter(self, labels: torch.Tensor) -> torch.Tensor:
        """encoder.py:111-113: drop timestamp / task tokens from the CTC targets when configured"""
        if self.config.remove_timestamps_from_ctc:
            labels = torch.nn.utils.rnn.pad_sequence([lab[lab < self.first_task_token] for lab in labels],
                                                     padding_value=-100).T
        return labels

    # ---- weight preparation ----------------------------------------------------------------------------------
    def _cache_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate_cache(self) -> None:
        self._prepared = None
        self.__dict__["_prepared_epoch"] = None
        _PREPARE_GRAPHS.pop(self, None)

    def prepare(self) -> dict:
        """bf16 / fused copies of the weights for the kernels; rebuilt when a parameter tensor changed."""
        ep = ops.prepare_epoch
        if ep and self._prepared is not None and self.__dict__.get("_prepared_epoch") == ep:
            return self._prepared  # same training step (see ops.prepare_epoch)
        key = self._cache_key()
        if self._prepared is not None and key == self._prepared_key:
            self.__dict__["_prepared_epoch"] = ep
            return self._prepared
        # Rebuild.  When the same parameter tensors merely carry new values (an optimizer step: same data_ptrs, new versions)
        # the rebuild is ~400 small cast / cat / stack launches issued from Python -- 3 ms of host time at the start of every
        # training step with the device idle.  From the second such rebuild on, the sequence is captured once into a CUDA graph
        # (its outputs live in the graph's memory pool) and replayed: one launch, the prepared tensors are refreshed in place.
        ptrs = tuple(k[0] for k in key)
        st = _PREPARE_GRAPHS.get(self)
        if st is not None and st["ptrs"] == ptrs and self._prepared is st["w"]:
            st["graph"].replay()
            self._prepared_key = key
            self.__dict__["_prepared_epoch"] = ep
            return self._prepared
        _PREPARE_GRAPHS.pop(self, None)
        same_tensors = self._prepared is not None and self._prepared_key is not None and \
            tuple(k[0] for k in self._prepared_key) == ptrs
        dev = self.conv1.weight.device
        if same_tensors and dev.type == "cuda" and prepare_graphs and not torch.cuda.is_current_stream_capturing():
            self._prepared = None  # release the eager copies before the graph's pool takes theirs
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                w = self._build_prepared()
            graph.replay()  # capture records, replay computes
            _PREPARE_GRAPHS[self] = {"ptrs": ptrs, "graph": graph, "w": w}
        else:
            w = self._build_prepared()
        self._prepared, self._prepared_key = w, key
        self.__dict__["_prepared_epoch"] = ep
        return w

    def _build_prepared(self) -> dict:
        cfg = self.config
        dev = self.conv1.weight.device
        d = cfg.d_model
        w: dict = {}
        w["conv1_w"], w["conv1_b"] = _conv_weight(self.conv1.weight), _f32(self.conv1.bias)
        w["conv2_w"], w["conv2_b"] = _conv_weight(self.conv2.weight), _f32(self.conv2.bias)
        w["pos"] = _f32(self.embed_positions.weight)
        full = cfg.use_fddt and not cfg.fddt_is_diagonal and not cfg.fddt_bias_only  # CustomLinear per class
        w["fddt_full"] = full
        if cfg.use_fddt and cfg.use_pre_pos_fddt and not full:
            w["fddt0"] = self.initial_fddt.tables()
        else:  # identity FDDT for the conv2 epilogue
            w["fddt0"] = (torch.ones(4, d, device=dev), torch.zeros(4, d, device=dev))
        if full:
            w["fddt"] = [f.full_tables() for f in self.fddts]
            w["fddt0_full"] = self.initial_fddt.full_tables() if cfg.use_pre_pos_fddt else None
        else:
            w["fddt"] = [f.tables() for f in self.fddts] if cfg.use_fddt else []
        layers = []
        for lyr in self.layers:
            e = _prep_attention(lyr.self_attn)
            e["ln1_g"], e["ln1_b"] = _f32(lyr.self_attn_layer_norm.weight), _f32(lyr.self_attn_layer_norm.bias)
            e["ln2_g"], e["ln2_b"] = _f32(lyr.final_layer_norm.weight), _f32(lyr.final_layer_norm.bias)
            e["w1"], e["b1"] = _bf16(lyr.fc1.weight), _f32(lyr.fc1.bias)
            e["w2"], e["b2"] = _bf16(lyr.fc2.weight), _f32(lyr.fc2.bias)
            layers.append(e)
        w["layers"] = layers
        w["lnf_g"], w["lnf_b"] = _f32(self.layer_norm.weight), _f32(self.layer_norm.bias)
        if hasattr(self, "ca_enrolls"):
            scbs = []
            for blk in self.ca_enrolls:
                e = _prep_attention(blk.cae.cross_attn, fuse_qkv=False)
                e["w1"], e["b1"] = _bf16(blk.cae.ffn[0].weight), _f32(blk.cae.ffn[0].bias)
                e["w2"], e["b2"] = _bf16(blk.cae.ffn[3].weight), _f32(blk.cae.ffn[3].bias)
                e["gate"] = _f32(blk.cae.cross_gate.gate)
                scbs.append(e)
            w["scb"] = scbs
        if hasattr(self, "additional_layer"):  # used INSTEAD of the extra self-attention (encoder.py:88-100)
            lyr = self.additional_layer
            e = _prep_attention(lyr.self_attn)
            e["ln1_g"], e["ln1_b"] = _f32(lyr.self_attn_layer_norm.weight), _f32(lyr.self_attn_layer_norm.bias)
            e["ln2_g"], e["ln2_b"] = _f32(lyr.final_layer_norm.weight), _f32(lyr.final_layer_norm.bias)
            e["w1"], e["b1"] = _bf16(lyr.fc1.weight), _f32(lyr.fc1.bias)
            e["w2"], e["b2"] = _bf16(lyr.fc2.weight), _f32(lyr.fc2.bias)
            w["ctc_layer"] = e
        elif hasattr(self, "additional_self_attention_layer"):
            w["ctc_attn"] = _prep_attention(self.additional_self_attention_layer)
        if hasattr(self, "subsample_conv1"):
            w["sub1"] = _conv_weight(self.subsample_conv1.weight)
            w["sub2"] = _conv_weight(self.subsample_conv2.weight)
        if hasattr(self, "lm_head"):
            w["lm_head"] = _bf16(self.lm_head.weight)
        return w

    # ---- kernels sequences -----------------------------------------------------------------------------------
    def _self_attention(self, e: dict, xin_bf16: torch.Tensor, Bx: int, T: int, out: torch.Tensor,
                        o_row_stride: int, o_batch_stride: int) -> None:
        """fused QKV projection + attention; writes bf16 context into ``out`` (pointer/strides given)."""
        d, H = self.config.d_model, self.config.encoder_attention_heads
        qkv = torch.empty(Bx * T, 3 * d, dtype=torch.bfloat16, device=xin_bf16.device)
        ops.gemm(xin_bf16, e["wqkv"], qkv, epilogue=ops.EPI_BIAS_BF16, bias=e["bqkv"])
        ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], out, B=Bx, H=H, Tq=T, Tk=T, q_row_stride=3 * d,
                      q_batch_stride=T * 3 * d, kv_row_stride=3 * d, kv_batch_stride=T * 3 * d,
                      o_row_stride=o_row_stride, o_batch_stride=o_batch_stride, variant=self.attention_variant)

    def _scb(self, e: dict, x: torch.Tensor, xb: torch.Tensor, B: int, T: int, kv: Optional[torch.Tensor] = None
             ) -> torch.Tensor:
        """SE-DiCoW speaker communication block (layers.py:145-170).  Default: the interleaved [2B, T, d] streams --
        x (fp32 residual, target rows updated in place), xb = bf16 copy of x.  With ``kv`` (bf16 [B, T, 2d], the
        enrollment stream's projected keys / values of this block from an earlier window of the same recording) x / xb
        hold the B target streams only.  Returns kv."""
        d, H = self.config.d_model, self.config.encoder_attention_heads
        dev = x.device
        sb = (1 if kv is not None else 2) * T * d  # batch stride of the target rows inside x / xb
        q = torch.empty(B, T, d, dtype=torch.bfloat16, device=dev)
        # q from the target rows (even), k/v from the enrollment rows (odd) -- no LayerNorm
        ops.gemm(xb, e["wq"], q, epilogue=ops.EPI_BIAS_BF16, bias=e["bq"], nb=B, Mb=T, lda=d,
                 a_batch_stride=sb, ldo=d, out_batch_stride=T * d)
        if kv is None:
            kv = torch.empty(B, T, 2 * d, dtype=torch.bfloat16, device=dev)
            ops.gemm(xb[1], e["wkv"], kv, epilogue=ops.EPI_BIAS_BF16, bias=e["bkv"], nb=B, Mb=T, lda=d,
                     a_batch_stride=2 * T * d, ldo=2 * d, out_batch_stride=T * 2 * d)
        ctx = torch.empty(B, T, d, dtype=torch.bfloat16, device=dev)
        ops.attention(q, kv, kv[:, :, d:], ctx, B=B, H=H, Tq=T, Tk=T, q_row_stride=d, q_batch_stride=T * d,
                      kv_row_stride=2 * d, kv_batch_stride=T * 2 * d, o_row_stride=d, o_batch_stride=T * d,
                      variant=self.attention_variant)
        ao = torch.empty(B, T, d, dtype=torch.bfloat16, device=dev)
        ops.gemm(ctx, e["wo"], ao, epilogue=ops.EPI_BIAS_BF16, bias=e["bo"])
        # ffn.0 over cat([attn_out, q_stream]) without materialising the concat: split-K over two sources
        hdn = torch.empty(B, T, e["w1"].shape[0], dtype=torch.bfloat16, device=dev)
        ops.gemm(ao, e["w1"], hdn, epilogue=ops.EPI_BIAS_GELU_BF16, bias=e["b1"], nb=B, Mb=T, K=2 * d, lda=d,
                 a_batch_stride=T * d, A2=xb, lda2=d, a2_batch_stride=sb, K1=d, ldo=hdn.shape[-1],
                 out_batch_stride=T * hdn.shape[-1])
        # target rows: x += tanh(gate) * (ffn.3(hdn))
        ops.gemm(hdn, e["w2"], x, epilogue=ops.EPI_RESIDUAL_F32, bias=e["b2"], nb=B, Mb=T, lda=hdn.shape[-1],
                 a_batch_stride=T * hdn.shape[-1], ldo=d, out_batch_stride=sb, resid=x, ldr=d,
                 resid_batch_stride=sb, gate=e["gate"])
        return kv

    def possibly_update_last_hidden_states(self, hidden_states: torch.Tensor) -> torch.Tensor:
        """encoder.py:87-106: extra self-attention (replaces the hidden state), two stride-2 convs.  fp32 in/out."""
        B, T, _ = hidden_states.shape
        return self._ctc_neck(self.prepare(), ops.cast_bf16(hidden_states), B, T).float()

    def _ctc_neck(self, w: dict, hb: torch.Tensor, B: int, T: int) -> torch.Tensor:
        """bf16 [B, T, d] -> bf16 [B, T', d] (T' = T/4 with sub-sampling)."""
        cfg = self.config
        d = cfg.d_model
        dev = hb.device
        sub = "sub1" in w
        if "ctc_layer" in w:  # a whole pre-LN encoder layer on the final hidden state (encoder.py:88-93)
            e = w["ctc_layer"]
            rows, ffn = B * T, e["w1"].shape[0]
            x = hb.view(rows, d).float()
            ln = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            ops.fddt_layernorm(x, gamma=e["ln1_g"], beta=e["ln1_b"], ln_out_bf16=ln)
            ctx = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            self._self_attention(e, ln, B, T, ctx, d, T * d)
            d1 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            ops.gemm(ctx, e["wo"], d1, epilogue=ops.EPI_BIAS_BF16, bias=e["bo"])
            ops.fddt_layernorm(x, gamma=e["ln2_g"], beta=e["ln2_b"], ln_out_bf16=ln, delta1=d1, store_x=False)
            hdn = torch.empty(rows, ffn, dtype=torch.bfloat16, device=dev)
            ops.gemm(ln, e["w1"], hdn, epilogue=ops.EPI_BIAS_GELU_BF16, bias=e["b1"])
            d2 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            ops.gemm(hdn, e["w2"], d2, epilogue=ops.EPI_BIAS_BF16, bias=e["b2"])
            hb = torch.empty(B, T, d, dtype=torch.bfloat16, device=dev)
            ops.fddt_layernorm(x, x_out_bf16=hb.view(rows, d), delta1=d1, delta2=d2, store_x=False)  # x + d1 + d2 -> bf16
            if not sub:
                return hb
            buf = torch.zeros(B, T + 2, d, dtype=torch.bfloat16, device=dev)
            buf[:, 1:T + 1] = hb
        elif "ctc_attn" in w:
            e = w["ctc_attn"]
            ctx = torch.empty(B * T, d, dtype=torch.bfloat16, device=dev)
            self._self_attention(e, hb.view(B * T, d), B, T, ctx, d, T * d)
            if sub:  # out_proj writes straight into the zero-padded channels-last buffer of the first conv
                buf = torch.empty(B, T + 2, d, dtype=torch.bfloat16, device=dev)
                ops.zero_pad_rows(buf)
                ops.gemm(ctx, e["wo"], buf[:, 1:], epilogue=ops.EPI_BIAS_BF16, bias=e["bo"], nb=B, Mb=T, lda=d,
                         a_batch_stride=T * d, ldo=d, out_batch_stride=(T + 2) * d)
            else:
                buf = torch.empty(B, T, d, dtype=torch.bfloat16, device=dev)
                ops.gemm(ctx, e["wo"], buf, epilogue=ops.EPI_BIAS_BF16, bias=e["bo"])
                return buf
        elif sub:
            buf = torch.zeros(B, T + 2, d, dtype=torch.bfloat16, device=dev)
            buf[:, 1:T + 1] = hb.view(B, T, d)
        else:
            return hb.view(B, T, d)
        T1 = (T + 2 - 3) // 2 + 1
        buf1 = torch.empty(B, T1 + 2, d, dtype=torch.bfloat16, device=dev)
        ops.zero_pad_rows(buf1)
        ops.gemm(buf, w["sub1"], buf1[:, 1:], epilogue=ops.EPI_BIAS_BF16, nb=B, Mb=T1, K=3 * d, lda=2 * d,
                 a_batch_stride=(T + 2) * d, ldo=d, out_batch_stride=(T1 + 2) * d)
        T2 = (T1 + 2 - 3) // 2 + 1
        out = torch.empty(B, T2, d, dtype=torch.bfloat16, device=dev)
        ops.gemm(buf1, w["sub2"], out, epilogue=ops.EPI_BIAS_BF16, nb=B, Mb=T2, K=3 * d, lda=2 * d,
                 a_batch_stride=(T1 + 2) * d, ldo=d, out_batch_stride=T2 * d)
        return out

    def ctc_logits_from_hidden(self, hidden_bf16: torch.Tensor, B: int, T: int, return_neck: bool = False):
        w = self.prepare()
        neck = self._ctc_neck(w, hidden_bf16, B, T)
        Tp = neck.shape[1]
        V1 = w["lm_head"].shape[0]
        # rows of the buffer padded to a multiple of 4 floats (V + 1 is odd for Whisper vocabularies): 16-byte stores in the GEMM
        # epilogue; the result is the [B, T', V + 1] view of it (unit column stride, not contiguous)
        ldl = -(-V1 // 4) * 4
        logits = torch.empty(B, Tp, ldl, dtype=torch.float32, device=neck.device)[..., :V1]
        ops.gemm(neck.view(B * Tp, -1), w["lm_head"], logits.as_strided((B * Tp, V1), (ldl, 1)), epilogue=ops.EPI_BIAS_F32)
        return (logits, neck) if return_neck else logits

    def forward(self, input_features, attention_mask=None, head_mask=None, output_attentions=None,
                output_hidden_states=None, return_dict=None, stno_mask=None, return_logits=False, enrollments=None,
                enrollment_kv=None, capture_enrollment_kv=None):
        """encoder.py:140-246.  With autograd recording, trainable parameters and ``return_logits`` (the CTC pre-training
        call, src/utils/trainers.py:76-103) the logits carry a grad_fn (training.EncoderLogitsFn); every other call is
        the inference path.  Fine-tuning goes through DiCoWForConditionalGeneration.forward, which owns the encoder's
        backward."""
        if return_logits and input_features.is_cuda:
            from . import training
            if training.trainable(self):
                if output_attentions or output_hidden_states or head_mask is not None:
                    raise NotImplementedError("output_attentions / output_hidden_states / head_mask are not produced by "
                                              "the fused B200 path")
                params = [p for p in self.parameters() if p.requires_grad]
                enr = enrollments or {}
                logits, hidden, neck = training.EncoderLogitsFn.apply(self, input_features, stno_mask,
                                                                      enr.get("input_features"), enr.get("stno_mask"), *params)
                return DiCoWCTCOutput(loss=None, logits=logits, hidden_states=neck, encoder_last_hidden_state=hidden)
        with torch.no_grad():
            return self._forward_inference(input_features, attention_mask, head_mask, output_attentions,
                                           output_hidden_states, return_dict, stno_mask, return_logits, enrollments,
                                           enrollment_kv, capture_enrollment_kv)

    def _forward_inference(self, input_features, attention_mask=None, head_mask=None, output_attentions=None,
                           output_hidden_states=None, return_dict=None, stno_mask=None, return_logits=False,
                           enrollments=None, enrollment_kv=None, capture_enrollment_kv=None):
        """``capture_enrollment_kv`` (a list) receives, per speaker communication block, the enrollment stream's projected
        keys / values (bf16 [B, T, 2d]); ``enrollment_kv`` takes such a list back instead of ``enrollments``: the
        enrollment stream never reads the target stream (layers.py:145-170 updates the query stream only), so for the
        next windows of the same recording its 8 layer passes and stem are skipped -- same numbers, 16 % fewer FLOPs per
        long-form window (generate() does this)."""
        cfg = self.config
        if enrollment_kv is not None and enrollments is not None:
            raise ValueError("pass either enrollments or enrollment_kv")
        if output_attentions or output_hidden_states or head_mask is not None:
            raise NotImplementedError("output_attentions / output_hidden_states / head_mask are not produced by the "
                                      "fused B200 path")
        if enrollments is not None:  # encoder.py:152-154
            input_features = torch.stack((input_features, enrollments["input_features"]), dim=1).flatten(0, 1)
            stno_mask = torch.stack((stno_mask, enrollments["stno_mask"]), dim=1).flatten(0, 1)
        expected = self.get_max_len()
        if input_features.shape[-1] != expected:  # encoder.py:156-160
            raise ValueError(f"Whisper expects the mel input features to be of length {expected}, but found "
                             f"{input_features.shape[-1]}. Make sure to pad the input mel features to {expected}.")
        if not input_features.is_cuda:
            raise ops.DicowError("DiCoWEncoder.forward needs CUDA tensors on an sm_100 device (no CPU fallback)")
        w = self.prepare()
        dev = input_features.device
        d, F = cfg.d_model, input_features.shape[-1]
        Bx, T = input_features.shape[0], F // 2
        feats = input_features.float().contiguous()
        if cfg.use_fddt:
            if stno_mask is None:
                raise ValueError("stno_mask is required when use_fddt is set")
            stno = stno_mask.to(device=dev, dtype=torch.float32).contiguous()
        else:
            stno = None
        # ---- stem: conv1+GELU, conv2+GELU, initial FDDT, + positions (encoder.py:167-179) ----
        a0 = torch.empty(Bx, F + 2, cfg.num_mel_bins, dtype=torch.bfloat16, device=dev)
        ops.features_to_channels_last(feats, a0)
        a1 = torch.empty(Bx, F + 2, d, dtype=torch.bfloat16, device=dev)
        ops.zero_pad_rows(a1)
        C = cfg.num_mel_bins
        ops.gemm(a0, w["conv1_w"], a1[:, 1:], epilogue=ops.EPI_BIAS_GELU_BF16, bias=w["conv1_b"], nb=Bx, Mb=F,
                 K=3 * C, lda=C, a_batch_stride=(F + 2) * C, ldo=d, out_batch_stride=(F + 2) * d)
        x = torch.empty(Bx, T, d, dtype=torch.float32, device=dev)
        if stno is not None and cfg.use_pre_pos_fddt:
            stno0 = stno
        else:  # identity transform: class 0 weight 1
            stno0 = torch.zeros(Bx, 4, T, dtype=torch.float32, device=dev)
            stno0[:, 0] = 1.0
        fw0, fb0 = w["fddt0"]
        if w["fddt_full"] and w.get("fddt0_full") is not None:
            # full-matrix initial FDDT (encoder.py:173-179 with CustomLinear): gelu(conv2) in bf16, the four class
            # transforms as one GEMM, mask-weighted sum + positions in fp32
            g2 = torch.empty(Bx * T, d, dtype=torch.bfloat16, device=dev)
            ops.gemm(a1, w["conv2_w"], g2, epilogue=ops.EPI_BIAS_GELU_BF16, bias=w["conv2_b"], nb=Bx, Mb=T, K=3 * d,
                     lda=2 * d, a_batch_stride=(F + 2) * d, ldo=d, out_batch_stride=T * d)
            W4, b4 = w["fddt0_full"]
            y = torch.empty(Bx * T, 4 * d, dtype=torch.bfloat16, device=dev)
            ops.gemm(g2, W4, y, epilogue=ops.EPI_BIAS_BF16, bias=b4)
            ops.fddt_full_combine(y, stno, x, T=T, pos=w["pos"])
            del g2, y
        else:
            ops.gemm(a1, w["conv2_w"], x, epilogue=ops.EPI_GELU_FDDT_POS_F32, bias=w["conv2_b"], nb=Bx, Mb=T, K=3 * d,
                     lda=2 * d, a_batch_stride=(F + 2) * d, ldo=d, out_batch_stride=T * d, stno=stno0,
                     stno_batch_stride=4 * T, fddt_w=fw0, fddt_b=fb0, pos=w["pos"])
        del a0, a1
        # ---- layers (encoder.py:191-223) ----
        # The out_proj / fc2 GEMMs write their bf16 outputs (d1, d2) instead of read-modify-writing the fp32 residual
        # stream in their epilogues; the pending updates are folded into the next FDDT+LayerNorm kernel:
        #   LN2 input  = x + d1              (x itself is not rewritten there)
        #   next layer : x <- FDDT(x + d1 + d2), LN1(x)
        # which is the reference's bf16-autocast arithmetic (bf16 Linear output + fp32 residual) with 30 % less HBM traffic.
        n_scb = cfg.scb_layers if (cfg.use_enrollments and cfg.scb_layers) else 0
        cached = enrollment_kv is not None
        if cached and len(enrollment_kv) != n_scb:
            raise ValueError(f"enrollment_kv needs one entry per speaker communication block ({n_scb})")
        if n_scb and not cached and (Bx % 2):
            raise ValueError("use_enrollments expects interleaved target/enrollment streams (even batch)")
        ffn = cfg.encoder_ffn_dim
        d1 = d2 = None  # pending bf16 residual updates of the previous layer
        for i, e in enumerate(w["layers"]):
            rows = Bx * T
            ln = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            fd = w["fddt"][i] if (cfg.use_fddt and i < len(w["fddt"])) else None
            if fd is not None and w["fddt_full"]:
                # full-matrix FDDT (layers.py:7-47): fold the pending deltas, project the bf16 stream with the stacked
                # [4 d, d] class transforms, take the mask-weighted sum back into the fp32 stream
                xb = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
                ops.fddt_layernorm(x, x_out_bf16=xb, delta1=d1, delta2=d2, store_x=True)
                y = torch.empty(rows, 4 * d, dtype=torch.bfloat16, device=dev)
                ops.gemm(xb, fd[0], y, epilogue=ops.EPI_BIAS_BF16, bias=fd[1])
                ops.fddt_full_combine(y, stno, x.view(rows, d), T=T)
                del xb, y
                fd, d1, d2 = None, None, None
            if i < n_scb:
                xb = torch.empty(Bx, T, d, dtype=torch.bfloat16, device=dev)
                ops.fddt_layernorm(x, T=T, stno=stno if fd else None, fddt_w=fd[0] if fd else None,
                                   fddt_b=fd[1] if fd else None, x_out_bf16=xb, delta1=d1, delta2=d2)
                d1 = d2 = None
                kv = self._scb(w["scb"][i], x, xb, Bx if cached else Bx // 2, T, kv=enrollment_kv[i] if cached else None)
                if capture_enrollment_kv is not None and not cached:
                    capture_enrollment_kv.append(kv)
                if i == n_scb - 1 and not cached:  # encoder.py:210-213: the enrollment stream is no longer needed
                    x = x.view(Bx // 2, 2, T, d)[:, 0].contiguous()
                    stno = stno.view(Bx // 2, 2, 4, T)[:, 0].contiguous() if stno is not None else None
                    Bx //= 2
                    rows = Bx * T
                    ln = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
                ops.fddt_layernorm(x, gamma=e["ln1_g"], beta=e["ln1_b"], ln_out_bf16=ln)
            else:
                ops.fddt_layernorm(x, T=T, stno=stno if fd else None, fddt_w=fd[0] if fd else None,
                                   fddt_b=fd[1] if fd else None, gamma=e["ln1_g"], beta=e["ln1_b"], ln_out_bf16=ln,
                                   delta1=d1, delta2=d2)
            ctx = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            self._self_attention(e, ln, Bx, T, ctx, d, T * d)
            d1 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            ops.gemm(ctx, e["wo"], d1, epilogue=ops.EPI_BIAS_BF16, bias=e["bo"])
            ops.fddt_layernorm(x, gamma=e["ln2_g"], beta=e["ln2_b"], ln_out_bf16=ln, delta1=d1, store_x=False)
            hdn = torch.empty(rows, ffn, dtype=torch.bfloat16, device=dev)
            ops.gemm(ln, e["w1"], hdn, epilogue=ops.EPI_BIAS_GELU_BF16, bias=e["b1"])
            d2 = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
            ops.gemm(hdn, e["w2"], d2, epilogue=ops.EPI_BIAS_BF16, bias=e["b2"])
            del hdn, ctx, ln
        # ---- final LayerNorm (encoder.py:228) ----
        out = torch.empty(Bx, T, d, dtype=torch.float32, device=dev)
        out_bf16 = torch.empty(Bx, T, d, dtype=torch.bfloat16, device=dev) if return_logits else None
        ops.fddt_layernorm(x, gamma=w["lnf_g"], beta=w["lnf_b"], ln_out_f32=out, ln_out_bf16=out_bf16, delta1=d1,
                           delta2=d2, store_x=False)
        if return_logits:  # encoder.py:233-240
            logits, neck = self.ctc_logits_from_hidden(out_bf16, Bx, T, return_neck=True)
            return DiCoWCTCOutput(loss=None, logits=logits, hidden_states=neck.float(), encoder_last_hidden_state=out)
        if return_dict is False:
            return (out,)
        return BaseModelOutput(last_hidden_state=out)
