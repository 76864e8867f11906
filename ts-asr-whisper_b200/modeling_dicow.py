"""Host-side mirror of the reference's top-level model for the hot path:

    DiCoW                           src/models/dicow/modeling_dicow.py:146-221  (WhisperModel with a DiCoWEncoder)
    DiCoWForConditionalGeneration   src/models/dicow/modeling_dicow.py:224-357  (+ DiCoWGenerationMixin, generation.py)
    SoftLabelCreator                src/models/dicow/modeling_dicow.py:23-144

Same class names, ``forward`` / ``generate`` signatures, ``state_dict`` keys (``model.encoder.*``, ``model.decoder.*``,
``proj_out.weight`` tied to ``model.decoder.embed_tokens.weight``) and outputs.  The modules are parameter containers:
the arithmetic of the decoder (HF WhisperDecoder, HF:models/whisper/modeling_whisper.py:449-796), proj_out, both losses
and the greedy token loop with its logits processors runs in libdicow_b200.so (ops.py).  ``generate()`` owns its
long-form loop instead of overriding private HF hooks (the reference targets transformers 4.55 internals that no longer
exist) and mirrors HF WhisperGenerationMixin.generate (HF:models/whisper/generation_whisper.py:
383-968) + the reference's overrides (generation.py:73-118 STNO slicing, :121-149 forced init tokens, :415-534
_retrieve_segment) for the supported mode: greedy, forced language/task prompt, timestamps on, no temperature fallback,
no conditioning on previous text (the recipes' configs/decode/*_greedy.yaml).  No eager / CPU fallback.
"""
from __future__ import annotations

import os
import re
from decimal import ROUND_HALF_UP, Decimal
from typing import Dict, List, Optional

import math

import torch
from torch import nn
from transformers import GenerationConfig, PreTrainedModel
from transformers.modeling_outputs import BaseModelOutput, Seq2SeqLMOutput, Seq2SeqModelOutput

from . import ops
from .configuration import DiCoWConfig
from .modeling import AttentionParams, DiCoWEncoder, _bf16, _f32


# ----------------------------------------------------------------------------------------------------------------
# parameter containers (names == HF WhisperDecoder state_dict)
# ----------------------------------------------------------------------------------------------------------------
class DecoderLayerParams(nn.Module):
    def __init__(self, d: int, ffn: int):
        super().__init__()
        self.self_attn = AttentionParams(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = AttentionParams(d)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, ffn)
        self.fc2 = nn.Linear(ffn, d)
        self.final_layer_norm = nn.LayerNorm(d)


class DiCoWDecoder(nn.Module):
    """Parameters of HF WhisperDecoder (HF:modeling_whisper.py:650-689)."""

    def __init__(self, config: DiCoWConfig):
        super().__init__()
        d = config.d_model
        self.config = config
        self.embed_tokens = nn.Embedding(config.vocab_size, d, padding_idx=config.pad_token_id)
        self.embed_positions = nn.Embedding(config.max_target_positions, d)
        self.layers = nn.ModuleList([DecoderLayerParams(d, config.decoder_ffn_dim) for _ in range(config.decoder_layers)])
        self.layer_norm = nn.LayerNorm(d)


LORA_TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")  # src/models/containers.py:73 (decoder only)


def lora_of(lin: nn.Module):
    """(A [r, in], B [out, r], alpha / r) of a linear layer that carries a LoRA adapter, else None"""
    A = getattr(lin, "lora_A", None)
    return None if A is None else (A, lin.lora_B, float(lin.lora_scale))


def effective_weight(lin: nn.Module) -> torch.Tensor:
    """fp32 W + (alpha / r) B A: what a LoRA-adapted layer multiplies by (peft's merged weight); W itself without an adapter.
    The kernels consume the merged weight -- one GEMM per layer, exactly as without the adapter."""
    W = lin.weight.detach().float()
    lo = lora_of(lin)
    if lo is None:
        return W
    A, B, s = lo
    return torch.addmm(W, B.detach().float(), A.detach().float(), alpha=s)


_AUX_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


class _HostCopy:
    """A small device tensor on the host without stalling the launch queue: copied to pinned memory on a side stream that
    waits only for what was queued before this call; ``get()`` blocks on that copy alone."""

    def __init__(self, t: torch.Tensor):
        self.ev = None
        if not t.is_cuda:
            self.host = t
            return
        idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
        side = _AUX_STREAMS.get(idx)
        if side is None:
            side = _AUX_STREAMS[idx] = torch.cuda.Stream(device=t.device)
        self.host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        side.wait_stream(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(side):
            self.host.copy_(t, non_blocking=True)
            self.ev = torch.cuda.Event()
            self.ev.record(side)
        t.record_stream(side)

    def get(self) -> torch.Tensor:
        if self.ev is not None:
            self.ev.synchronize()
            self.ev = None
        return self.host


def shift_tokens_right(labels: torch.Tensor, pad_token_id: int, decoder_start_token_id: int) -> torch.Tensor:
    """HF:modeling_whisper.py shift_tokens_right (index bookkeeping on int64 labels; call site modeling_dicow.py:275-279)."""
    shifted = labels.new_zeros(labels.shape)
    shifted[:, 1:] = labels[:, :-1].clone()
    shifted[:, 0] = decoder_start_token_id
    shifted.masked_fill_(shifted == -100, pad_token_id)
    return shifted


class SoftLabelCreator(nn.Module):
    """Timestamp-smoothing table of src/models/dicow/modeling_dicow.py:23-70.  The reference stores a dense
    [num_ts, vocab] matrix and builds dense one-hot targets; the timestamp ids of a Whisper tokenizer are one contiguous
    block, so the kernel only needs ``ts_begin`` and the [num_ts, num_ts] Gaussian weights."""

    def __init__(self, tokenizer, timestamp_sigma: float = 0.08):
        super().__init__()
        self.tokenizer = tokenizer
        self.timestamp_sigma = timestamp_sigma
        pat = re.compile(r"<\|(\d+\.\d+)\|>")
        id_to_time = {}
        for tok, idx in tokenizer.get_vocab().items():
            m = pat.match(tok)
            if m:
                id_to_time[idx] = float(m.group(1))
        self.ts_begin, self.smoothing = 0, None
        if id_to_time:
            ids = sorted(id_to_time)
            if ids != list(range(ids[0], ids[0] + len(ids))):
                raise NotImplementedError("timestamp token ids are expected to be one contiguous block")
            times = torch.tensor([id_to_time[i] for i in ids])
            w = torch.exp(-((times[:, None] - times[None, :]) ** 2) / (2 * timestamp_sigma ** 2))
            self.ts_begin = ids[0]
            self.register_buffer("ts_smoothing_weights", (w / w.sum(dim=1, keepdim=True)).contiguous(), persistent=False)
            self.smoothing = True

    def smoothing_on(self, device) -> Optional[torch.Tensor]:
        """the [num_ts, num_ts] table on ``device``.  set_tokenizer() usually runs after the model has been moved, so the
        buffer sits on the host: copying 9 MB of pageable memory per step cost 0.66 ms and a stream synchronisation."""
        if not self.smoothing:
            return None
        t = self.ts_smoothing_weights
        if t.device == torch.device(device):
            return t
        cached = self.__dict__.get("_smoothing_dev")
        if cached is None or cached.device != torch.device(device):
            cached = t.to(device)
            self.__dict__["_smoothing_dev"] = cached
        return cached

    def compute_loss(self, logits: torch.Tensor, labels: torch.Tensor, upp_labels: Optional[torch.Tensor]) -> torch.Tensor:
        V = logits.shape[-1]
        sm = self.smoothing_on(logits.device)
        return ops.softlabel_ce(logits.reshape(-1, V), labels.to(logits.device),
                                upp_labels.to(logits.device) if upp_labels is not None else None,
                                ts_begin=self.ts_begin, smoothing=sm, soft_mode=True)


# ----------------------------------------------------------------------------------------------------------------
# base model
# ----------------------------------------------------------------------------------------------------------------
class DiCoW(nn.Module):
    """encoder + decoder (src/models/dicow/modeling_dicow.py:146-221)."""

    def __init__(self, config: DiCoWConfig):
        super().__init__()
        self.config = config
        self.encoder = DiCoWEncoder(config)
        self.decoder = DiCoWDecoder(config)
        self._prepared: Optional[dict] = None
        self._prepared_key = None
        self._prepared_gen = 0  # bumped every time the prepared weights are rebuilt (captured graphs compare it)

    def invalidate_cache(self) -> None:
        """drop the prepared bf16 weights of the decoder AND the encoder (needed after in-place updates that do not bump
        ``Parameter._version``: ``p.data.copy_()``, EMA swaps, some sharded-optimizer paths)"""
        self._prepared = None
        self.__dict__["_prepared_epoch"] = None
        self.encoder.invalidate_cache()

    def get_encoder(self):
        return self.encoder

    def get_decoder(self):
        return self.decoder

    # ---- bf16 / fused decoder weights ------------------------------------------------------------------------
    def prepare_decoder(self) -> dict:
        ep = ops.prepare_epoch
        if ep and self._prepared is not None and self.__dict__.get("_prepared_epoch") == ep:
            return self._prepared  # same training step (see ops.prepare_epoch)
        key = tuple((p.data_ptr(), p._version) for p in self.decoder.parameters())
        if self._prepared is not None and key == self._prepared_key:
            self.__dict__["_prepared_epoch"] = ep
            return self._prepared
        dec = self.decoder
        sc = 64 ** -0.5  # folded q scale (HF:modeling_whisper.py:310), exact in bf16

        def att(a: AttentionParams) -> dict:
            d = a.q_proj.weight.shape[0]
            zeros = torch.zeros(d, device=a.q_proj.weight.device)
            return {"wq": _bf16(effective_weight(a.q_proj) * sc), "bq": (a.q_proj.bias.detach().float() * sc).contiguous(),
                    "wkv": _bf16(torch.cat([effective_weight(a.k_proj), effective_weight(a.v_proj)], 0)),
                    "bkv": torch.cat([zeros, a.v_proj.bias.detach().float()]).contiguous(),
                    "wo": _bf16(effective_weight(a.out_proj)), "bo": _f32(a.out_proj.bias)}

        w: dict = {"tok": _f32(dec.embed_tokens.weight), "pos": _f32(dec.embed_positions.weight),
                   "proj": _bf16(dec.embed_tokens.weight),  # proj_out is tied to embed_tokens (train.py:109-113)
                   "lnf_g": _f32(dec.layer_norm.weight), "lnf_b": _f32(dec.layer_norm.bias), "layers": []}
        for lyr in dec.layers:
            e = {"self": att(lyr.self_attn), "cross": att(lyr.encoder_attn)}
            s = e["self"]
            s["wqkv"] = torch.cat([s["wq"], s["wkv"]], 0).contiguous()
            s["bqkv"] = torch.cat([s["bq"], s["bkv"]]).contiguous()
            for nm, ln in (("ln1", lyr.self_attn_layer_norm), ("ln2", lyr.encoder_attn_layer_norm),
                           ("ln3", lyr.final_layer_norm)):
                e[nm + "_g"], e[nm + "_b"] = _f32(ln.weight), _f32(ln.bias)
            e["w1"], e["b1"] = _bf16(effective_weight(lyr.fc1)), _f32(lyr.fc1.bias)
            e["w2"], e["b2"] = _bf16(effective_weight(lyr.fc2)), _f32(lyr.fc2.bias)
            w["layers"].append(e)
        self._prepared, self._prepared_key = w, key
        self.__dict__["_prepared_epoch"] = ep
        self._prepared_gen += 1
        w["generation"] = self._prepared_gen
        return w

    # ---- teacher-forced decoder (training / evaluation forward) -------------------------------------------------
    @torch.no_grad()
    def decode_teacher_forced(self, decoder_input_ids: torch.Tensor, enc_bf16: torch.Tensor):
        """[B, S] ids + encoder states bf16 [B, T, d] -> (hidden fp32 [B, S, d], hidden bf16 [B*S, d]).
        HF WhisperDecoder.forward without cache (HF:modeling_whisper.py:691-796): embed + learned positions,
        layers x [causal self-attn, cross-attn, MLP] (pre-LN), final LN."""
        cfg = self.config
        w = self.prepare_decoder()
        dev = enc_bf16.device
        B, S = decoder_input_ids.shape
        T, d, H = enc_bf16.shape[1], cfg.d_model, cfg.decoder_attention_heads
        if S > cfg.max_target_positions:
            raise ValueError(f"decoder sequence length {S} exceeds max_target_positions {cfg.max_target_positions}")
        ids = decoder_input_ids.to(device=dev, dtype=torch.int64).contiguous()
        x = torch.empty(B, S, d, dtype=torch.float32, device=dev)
        ops.embed_tokens(ids, w["tok"], w["pos"], x, S=S, past=0)
        rows = B * S
        xf = x.view(rows, d)
        ln = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        ctx = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        qkv = torch.empty(rows, 3 * d, dtype=torch.bfloat16, device=dev)
        q = torch.empty(rows, d, dtype=torch.bfloat16, device=dev)
        kv = torch.empty(B * T, 2 * d, dtype=torch.bfloat16, device=dev)
        hdn = torch.empty(rows, cfg.decoder_ffn_dim, dtype=torch.bfloat16, device=dev)
        encf = enc_bf16.reshape(B * T, d)
        for e in w["layers"]:
            s, c = e["self"], e["cross"]
            ops.fddt_layernorm(x, gamma=e["ln1_g"], beta=e["ln1_b"], ln_out_bf16=ln)
            ops.gemm(ln, s["wqkv"], qkv, epilogue=ops.EPI_BIAS_BF16, bias=s["bqkv"])
            ops.attention(qkv, qkv[:, d:], qkv[:, 2 * d:], ctx, B=B, H=H, Tq=S, Tk=S, q_row_stride=3 * d,
                          q_batch_stride=S * 3 * d, kv_row_stride=3 * d, kv_batch_stride=S * 3 * d, o_row_stride=d,
                          o_batch_stride=S * d, causal=True)
            ops.gemm(ctx, s["wo"], xf, epilogue=ops.EPI_RESIDUAL_F32, bias=s["bo"], resid=xf)
            ops.fddt_layernorm(x, gamma=e["ln2_g"], beta=e["ln2_b"], ln_out_bf16=ln)
            ops.gemm(ln, c["wq"], q, epilogue=ops.EPI_BIAS_BF16, bias=c["bq"])
            ops.gemm(encf, c["wkv"], kv, epilogue=ops.EPI_BIAS_BF16, bias=c["bkv"])
            ops.attention(q, kv, kv[:, d:], ctx, B=B, H=H, Tq=S, Tk=T, q_row_stride=d, q_batch_stride=S * d,
                          kv_row_stride=2 * d, kv_batch_stride=T * 2 * d, o_row_stride=d, o_batch_stride=S * d)
            ops.gemm(ctx, c["wo"], xf, epilogue=ops.EPI_RESIDUAL_F32, bias=c["bo"], resid=xf)
            ops.fddt_layernorm(x, gamma=e["ln3_g"], beta=e["ln3_b"], ln_out_bf16=ln)
            ops.gemm(ln, e["w1"], hdn, epilogue=ops.EPI_BIAS_GELU_BF16, bias=e["b1"])
            ops.gemm(hdn, e["w2"], xf, epilogue=ops.EPI_RESIDUAL_F32, bias=e["b2"], resid=xf)
        hid = torch.empty(B, S, d, dtype=torch.float32, device=dev)
        ops.fddt_layernorm(x, gamma=w["lnf_g"], beta=w["lnf_b"], ln_out_f32=hid, ln_out_bf16=ln)
        return hid, ln

    @torch.no_grad()
    def forward(self, input_features=None, attention_mask=None, stno_mask=None, decoder_input_ids=None,
                decoder_attention_mask=None, head_mask=None, decoder_head_mask=None, cross_attn_head_mask=None,
                encoder_outputs=None, past_key_values=None, decoder_inputs_embeds=None, decoder_position_ids=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None,
                cache_position=None, enrollments=None):
        if past_key_values is not None or decoder_inputs_embeds is not None or decoder_position_ids is not None:
            raise NotImplementedError("the B200 path decodes through generate(); HF cache objects / input embeddings "
                                      "are not accepted by forward()")
        if encoder_outputs is None:
            encoder_outputs = self.encoder(input_features, stno_mask=stno_mask, enrollments=enrollments)
        enc = encoder_outputs[0] if not isinstance(encoder_outputs, torch.Tensor) else encoder_outputs
        hid, _ = self.decode_teacher_forced(decoder_input_ids, ops.cast_bf16(enc.float()))
        return Seq2SeqModelOutput(last_hidden_state=hid, encoder_last_hidden_state=enc)


# ----------------------------------------------------------------------------------------------------------------
# greedy token loop
# ----------------------------------------------------------------------------------------------------------------
class _GreedyState:
    """Per-(device, batch size) buffers and the two captured CUDA graphs of one decoder step."""

    def __init__(self, model: "DiCoWForConditionalGeneration", B: int, T: int, dev: torch.device, beams: int = 1):
        """B = decoder rows (hypotheses); with ``beams`` > 1 the rows are utterance-major (row = u * beams + k), the
        cross-attention cache is per utterance (B / beams entries) and the beam-search state is allocated"""
        cfg = model.config
        d, L = cfg.d_model, cfg.decoder_layers
        self.B, self.T, self.beams = B, T, beams
        self.S_max = cfg.max_target_positions
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.ids = torch.zeros(B, self.S_max + 1, dtype=torch.int64, device=dev)
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.unfinished = torch.ones(B, dtype=torch.int32, device=dev)
        self.x = torch.empty(B, d, dtype=torch.float32, device=dev)
        self.ln = torch.empty(B, d, **bf)
        self.q = torch.empty(B, d, **bf)
        self.ctx = torch.empty(B, d, **bf)
        self.h = torch.empty(B, cfg.decoder_ffn_dim, **bf)
        self.self_kv = torch.zeros(L, B, self.S_max, 2 * d, **bf)
        # cross-attention cache, head-major [L, B, H, T, k(64) | v(64)]: one contiguous stream per (batch, head) and step
        U = B // beams
        self.cross_kv = torch.zeros(L, U, cfg.decoder_attention_heads, T, 128, **bf)  # (padding rows stay finite)
        self.cross_kv_rows = torch.empty(U * T, 2 * d, **bf)  # the projection GEMM's output before the re-layout
        self.ancestry = None
        if beams > 1:  # beam search: see DiCoWForConditionalGeneration.beam_decode_window
            i32, f32 = dict(dtype=torch.int32, device=dev), dict(dtype=torch.float32, device=dev)
            self.ancestry = torch.empty(B, self.S_max, **i32)
            self.ancestry_tmp = torch.empty(B, self.S_max, **i32)
            self.run_score, self.fin_score = torch.empty(B, **f32), torch.empty(B, **f32)
            self.fin_flag, self.unsat = torch.empty(B, **i32), torch.empty(U, **i32)
            self.fin_ids = torch.empty(B, self.S_max + 1, dtype=torch.int64, device=dev)
            self.ids_tmp = torch.empty(2 * B, self.S_max + 1, dtype=torch.int64, device=dev)
            self.scratch_i32, self.scratch_f32 = torch.empty(3 * B, **i32), torch.empty(2 * B, **f32)
            self.flags = torch.zeros(U, 4, **i32)
            self.cand = None
            self.ctc_r_tmp = None
        self.logits = torch.empty(B, cfg.vocab_size, dtype=torch.float32, device=dev)
        self.ctc = None      # ops.CtcJointState of joint CTC / attention decoding (allocated on first use)
        self.ctc_key = None
        self.proc = None     # [B, V] processed scores handed from the rules kernel to the joint CTC step
        self.graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        # the prepared-weight dict the graphs were captured with: a STRONG reference compared with ``is`` (an id() of a
        # freed dict can be handed to a later one, and the graphs would replay with pointers to freed weights)
        self.weights = None
        self.ctc_gen = 0  # bumped when the joint-CTC state object is replaced (part of the graph key)
        # persistent decode-layers kernel (ops.decode_layers): arrival counter of its grid barrier, layer pointer table
        self.mega_bar = torch.zeros(1, dtype=torch.int64, device=dev)
        self.mega_ws = torch.zeros(B * cfg.decoder_attention_heads * 136, dtype=torch.float32, device=dev)
        self.mega_table = None
        self.mega_weights = None


_DECODE_BUCKETS = (1, 2, 4, 8, 12, 16, 24, 32, 48, 64)


def _bucket_rows(B: int) -> int:
    """decode-state batch sizes are bucketed: the long-form loop shrinks its batch as recordings finish, and one state
    (K/V caches, staging rows, captured graphs: ~47 MB x B for large-v3-turbo) per distinct size would pile up"""
    for b in _DECODE_BUCKETS:
        if B <= b:
            return b
    return B


class _DecodeCache:
    """LRU of _GreedyState objects (at most ``capacity``).  Evicted states drop their buffers and captured graphs.
    ``copy.deepcopy(model)`` yields an empty cache (CUDA graphs cannot be copied; they are re-captured on demand)."""

    def __init__(self, capacity: int = 3):
        self.capacity = capacity
        self.states: "Dict[tuple, _GreedyState]" = {}

    def get(self, key, make):
        st = self.states.pop(key, None)
        if st is None:
            while len(self.states) >= self.capacity:
                old = self.states.pop(next(iter(self.states)))
                old.graphs.clear()
            st = make()
        self.states[key] = st  # most recently used last
        return st

    def clear(self) -> None:
        for st in self.states.values():
            st.graphs.clear()
        self.states.clear()

    def __len__(self):
        return len(self.states)

    def __deepcopy__(self, memo):
        return _DecodeCache(self.capacity)


_SPECULATION_STREAMS: Dict[tuple, "torch.cuda.Stream"] = {}  # one side stream per device (module level: models stay copyable)


class DiCoWForConditionalGeneration(PreTrainedModel):
    config_class = DiCoWConfig
    base_model_prefix = "model"
    main_input_name = "input_features"
    _tied_weights_keys = {"proj_out.weight": "model.decoder.embed_tokens.weight"}
    # checkpoints written by the reference after set_tokenizer() carry its dense [num_ts, vocab] smoothing matrix as a
    # persistent buffer (modeling_dicow.py:33); here the table is rebuilt from the tokenizer and is not part of the state
    _keys_to_ignore_on_load_unexpected = [r"soft_label_creator\.ts_smoothing_matrix"]
    _no_split_modules = ["EncoderLayerParams", "DecoderLayerParams"]
    # step graphs are captured once per batch size; set False to launch the step kernels eagerly (debugging)
    use_cuda_graphs = True
    # True: fused q|k,v projection, cluster split-K linear layers, few-rows LayerNorm kernel; "ln_prologue": LayerNorm as
    # the prologue of the linear kernel instead (measured slower); False: one kernel per operation (_decode_step_unfused)
    fused_decode_step = True
    # EXPERIMENTAL, off by default: greedy steps of <= 16 rows run the decoder layers as ONE persistent kernel with grid
    # barriers between the phases (csrc/decode_mega.cu; DICOW_DECODE_MEGA=1 or decode_megakernel = True).  Token-parity green,
    # but measured SLOWER than the kernel-per-operation sequence at B = 16 (0.50 vs 0.42 ms per step): DESIGN.md section 4.2
    decode_megakernel = os.environ.get("DICOW_DECODE_MEGA", "0") == "1"

    def __init__(self, config: DiCoWConfig):
        super().__init__(config)
        self.model = DiCoW(config)
        self.proj_out = nn.Linear(config.d_model, config.vocab_size, bias=False)
        self.max_target_positions = config.max_target_positions
        self.encoder_logits = None
        self.tokenizer = None
        self.stno_mask = None
        self.stno_mask_seek = None
        self.soft_label_creator = None
        self._greedy = _DecodeCache()
        self.post_init()
        self.tie_weights()
        if getattr(self, "generation_config", None) is None:
            self.generation_config = GenerationConfig.from_model_config(config)

    # ---- HF plumbing ---------------------------------------------------------------------------------------
    def _init_weights(self, module):
        """HF WhisperPreTrainedModel._init_weights semantics for the container modules + the DiCoW modules' own inits
        (src/models/dicow/encoder.py:79-82).  from_pretrained() calls this on EVERY module after loading and marks the
        tensors that came from the checkpoint with ``_is_hf_initialized``: only unmarked tensors are touched (an unguarded
        init here would overwrite the loaded weights)."""
        std = getattr(self.config, "init_std", 0.02)
        from .modeling import FDDT, CrossAttentionEnrollBlock, CustomDiagonalLinear, CustomLinear, DiCoWEncoder, Gate

        def fresh(t) -> bool:
            return t is not None and not getattr(t, "_is_hf_initialized", False) and t.device.type != "meta"

        if isinstance(module, (CustomDiagonalLinear, CustomLinear)):
            module.reset_parameters(weight=fresh(module.weight), bias=fresh(module.bias))  # encoder.py:79-82
        elif isinstance(module, Gate):
            if all(fresh(p) for p in module.parameters(recurse=False)):
                module.reset_parameters()
        elif isinstance(module, CrossAttentionEnrollBlock):
            if fresh(module.ffn[0].weight) and fresh(module.ffn[3].weight):
                module.reset_parameters()  # layers.py:95-110
        elif isinstance(module, FDDT) and module.bias_only:
            for c in ("target", "non_target", "overlap", "silence"):
                p = getattr(module, c + "_linear", None)
                if fresh(p):
                    p.data.zero_()  # FDDT.py:10
        elif isinstance(module, DiCoWEncoder):  # HF WhisperPreTrainedModel._init_weights: sinusoidal positions
            pos = module.embed_positions.weight
            if fresh(pos):
                n, ch = pos.shape
                inc = math.log(10000.0) / (ch // 2 - 1)
                inv = torch.exp(-inc * torch.arange(ch // 2, dtype=torch.float32))
                ang = torch.arange(n, dtype=torch.float32)[:, None] * inv[None, :]
                pos.data.copy_(torch.cat([ang.sin(), ang.cos()], dim=1).to(pos.device))
        elif isinstance(module, (nn.Linear, nn.Conv1d)):
            if getattr(module, "_dicow_custom_init", False):
                return
            if fresh(module.weight):
                module.weight.data.normal_(mean=0.0, std=std)
            if fresh(module.bias):
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            if fresh(module.weight) and module is not getattr(getattr(self.model, "encoder", None), "embed_positions", None):
                module.weight.data.normal_(mean=0.0, std=std)
                if module.padding_idx is not None:
                    module.weight.data[module.padding_idx].zero_()
        elif isinstance(module, nn.LayerNorm):
            if fresh(module.weight):
                module.weight.data.fill_(1.0)
            if fresh(module.bias):
                module.bias.data.zero_()

    def tie_weights(self, missing_keys=None, **kwargs):
        """proj_out is tied to the decoder's token embedding (src/train.py:109-113).  A checkpoint stores the tensor
        once: the tied key is not "missing" (left in the set, from_pretrained() would re-initialise proj_out.weight --
        which IS the embedding -- and overwrite the loaded embeddings)."""
        self.proj_out.weight = self.model.decoder.embed_tokens.weight
        if missing_keys is not None:
            missing_keys.discard("proj_out.weight")

    def get_encoder(self):
        return self.model.get_encoder()

    def get_decoder(self):
        return self.model.get_decoder()

    def get_output_embeddings(self):
        return self.proj_out

    def set_output_embeddings(self, new_embeddings):
        self.proj_out = new_embeddings

    def get_input_embeddings(self):
        return self.model.decoder.embed_tokens

    def can_generate(self) -> bool:
        return True

    def freeze_encoder(self):
        for p in self.model.encoder.parameters():
            p.requires_grad_(False)

    @property
    def _ddp_params_and_buffers_to_ignore(self):
        """read by torch DistributedDataParallel when it wraps the model: see parallel.ddp_ignore_list (the per-layer
        gradient exchange overlapped with the backward replaces DDP's end-of-backward all-reduce)"""
        from . import parallel
        return parallel.ddp_ignore_list(self)

    # ---- LoRA on the decoder (src/models/containers.py:69-78) -----------------------------------------------------
    def add_lora(self, r: int = 16, lora_alpha: int = 32, target_modules=LORA_TARGETS, seed: Optional[int] = None):
        """What ``get_peft_model(model, LoraConfig(r=16, lora_alpha=32, target_modules=r".*decoder.*(q_proj|k_proj|v_proj|
        out_proj|fc1|fc2).*", lora_dropout=0.0, bias="none"))`` does to the reference model, without peft (not in this image):
        every targeted decoder projection gets ``lora_A`` [r, in] (kaiming-uniform, a = sqrt(5): peft's init) and ``lora_B``
        [out, r] (zeros), the base model is frozen and the adapters are trainable.  The kernels multiply by the merged weight
        W + (alpha / r) B A (prepare_decoder); the backward projects onto A and B with four r-wide GEMMs per layer
        (training._linear_backward).  Parameter names contain "lora_" as the reference's freezing loop expects."""
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        for p in self.parameters():
            p.requires_grad_(False)
        n = 0
        for name, mod in self.model.decoder.named_modules():
            if isinstance(mod, nn.Linear) and name.rsplit(".", 1)[-1] in target_modules and lora_of(mod) is None:
                A = torch.empty(r, mod.in_features)
                bound = 1.0 / math.sqrt(mod.in_features)  # kaiming_uniform_(a=sqrt(5)) on [r, in]: U(-1/sqrt(in), 1/sqrt(in))
                A.uniform_(-bound, bound, generator=gen)
                mod.register_parameter("lora_A", nn.Parameter(A.to(mod.weight.device)))
                mod.register_parameter("lora_B", nn.Parameter(torch.zeros(mod.out_features, r, device=mod.weight.device)))
                mod.lora_scale = float(lora_alpha) / float(r)
                mod.weight._dicow_lora = mod  # how the backward finds the adapter from the base weight it is handed
                n += 1
        self.model.invalidate_cache()
        self.clear_decode_cache()
        return n

    def lora_state_dict(self) -> dict:
        """the adapter in peft's saved-adapter key format (``base_model.model.<module path>.lora_A.weight``)"""
        out = {}
        for name, mod in self.named_modules():
            if lora_of(mod) is not None:
                out[f"base_model.model.{name}.lora_A.weight"] = mod.lora_A.detach().clone()
                out[f"base_model.model.{name}.lora_B.weight"] = mod.lora_B.detach().clone()
        return out

    def load_lora_state_dict(self, sd: dict) -> None:
        mods = dict(self.named_modules())
        with torch.no_grad():
            for k, v in sd.items():
                m = re.match(r"(?:base_model\.model\.)?(.*)\.lora_([AB])(?:\.default)?(?:\.weight)?$", k)
                if m is None or m.group(1) not in mods or lora_of(mods[m.group(1)]) is None:
                    raise KeyError(f"no LoRA adapter for {k}")
                getattr(mods[m.group(1)], "lora_" + m.group(2)).copy_(v)
        self.invalidate_prepared()

    def merge_lora(self) -> None:
        """fold the adapters into the base weights and drop them (peft's merge_and_unload)"""
        with torch.no_grad():
            for mod in self.modules():
                if lora_of(mod) is not None:
                    mod.weight.copy_(effective_weight(mod).to(mod.weight.dtype))
                    del mod._parameters["lora_A"], mod._parameters["lora_B"]
                    mod.weight._dicow_lora = None
        self.invalidate_prepared()

    def clear_decode_cache(self) -> None:
        """free the decode states (self / cross K/V caches, staging rows, captured CUDA graphs) kept between generate() calls"""
        self._greedy.clear()

    def invalidate_prepared(self) -> None:
        """drop every prepared bf16 weight copy and everything captured against them; call after updating parameters in a
        way that does not bump ``Parameter._version`` (``p.data.copy_()``, EMA weight swaps, offload / sharded optimizers)"""
        self.model.invalidate_cache()
        self.clear_decode_cache()

    def train(self, mode: bool = True):
        if mode:  # a training phase follows: the decode states (tens of MB per row) are dead weight until the next eval
            self.clear_decode_cache()
        return super().train(mode)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self.invalidate_prepared()
        return out

    # ---- reference API -------------------------------------------------------------------------------------
    def set_tokenizer(self, tokenizer):  # modeling_dicow.py:237-240
        self.tokenizer = tokenizer
        self.soft_label_creator = SoftLabelCreator(tokenizer)

    def get_enc_logits(self, hidden_states: torch.Tensor) -> torch.Tensor:  # modeling_dicow.py:242-246
        enc = self.model.get_encoder()
        B, T, _ = hidden_states.shape
        return enc.ctc_logits_from_hidden(ops.cast_bf16(hidden_states.float()), B, T)

    def _ctc_labels(self, labels: torch.Tensor) -> torch.Tensor:
        """CTC targets of the joint loss (modeling_dicow.py:326-333): the decoder labels minus the common prompt tokens,
        eos -> -100; int64 index bookkeeping."""
        cfg = self.config
        enc_labels = labels.clone()
        prefix = getattr(self.tokenizer, "prefix_tokens", None) if self.tokenizer is not None else None
        if prefix is None:
            prefix = getattr(self, "ctc_prefix_tokens", ())
        for tok in prefix:  # modeling_dicow.py:330-332
            if enc_labels.shape[1] and bool((enc_labels[:, 0] == tok).all()):
                enc_labels = enc_labels[:, 1:]
        enc_labels[enc_labels == cfg.eos_token_id] = -100
        return enc_labels

    def forward(self, input_features=None, attention_mask=None, stno_mask=None, decoder_input_ids=None,
                decoder_attention_mask=None, head_mask=None, decoder_head_mask=None, cross_attn_head_mask=None,
                encoder_outputs=None, past_key_values=None, decoder_inputs_embeds=None, decoder_position_ids=None,
                labels=None, upp_labels=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict=None, cache_position=None, forced_decoder_ids=None, enrollments=None):
        """src/models/dicow/modeling_dicow.py:248-354.  With labels, autograd recording and trainable parameters the loss
        carries a grad_fn whose backward is the hand-scheduled kernel sequence of training.DiCoWTrainStepFn (what HF
        Trainer.training_step -> loss.backward() runs); otherwise forward values only."""
        cfg = self.config
        if labels is not None and decoder_input_ids is None and decoder_inputs_embeds is None:
            decoder_input_ids = shift_tokens_right(labels, cfg.pad_token_id, cfg.decoder_start_token_id)
        if past_key_values is not None or decoder_inputs_embeds is not None:
            raise NotImplementedError("forward() is the teacher-forced path; token-by-token decoding is generate()")
        if output_attentions or output_hidden_states or any(m is not None for m in (head_mask, decoder_head_mask,
                                                                                    cross_attn_head_mask)):
            raise NotImplementedError("output_attentions / output_hidden_states / head masks are not produced by the fused "
                                      "B200 path")
        if labels is not None and encoder_outputs is None and input_features is not None and input_features.is_cuda:
            from . import training
            if training.trainable(self):
                enc_model = self.model.get_encoder()
                enc_labels = None
                if cfg.ctc_weight > 0.0:
                    # The CTC targets are data-dependent index bookkeeping (prefix tokens stripped only if the whole batch
                    # carries them, the label range check, the optional timestamp filter: modeling_dicow.py:326-333,
                    # encoder.py:109-113) -- each test is a device-to-host read.  Done here on the device labels they stalled
                    # the launch queue at the start of every step (the host waited for the previous step's optimizer kernels,
                    # then the device idled 2 ms while Python caught up).  The labels are copied to pinned memory on a side
                    # stream instead and the bookkeeping runs on that copy when the CTC loss needs it, after the encoder
                    # and decoder launches are queued.
                    host = _HostCopy(labels)

                    def enc_labels():
                        lab = self._ctc_labels(host.get())
                        if lab.numel() and lab.max() >= cfg.vocab_size:  # encoder.py:109-110
                            raise ValueError(f"Label values must be <= vocab_size: {cfg.vocab_size}")
                        return enc_model.ctc_label_filter(lab)
                params = [p for p in self.parameters() if p.requires_grad]
                enr = enrollments or {}
                loss, logits, enc = training.DiCoWTrainStepFn.apply(self, input_features, stno_mask, decoder_input_ids, labels,
                                                                    upp_labels, enc_labels, enr.get("input_features"),
                                                                    enr.get("stno_mask"), *params)
                if return_dict is False:
                    return (loss, logits, enc)
                return Seq2SeqLMOutput(loss=loss, logits=logits, encoder_last_hidden_state=enc)
        with torch.no_grad():
            return self._forward_inference(input_features, stno_mask, decoder_input_ids, encoder_outputs, labels, upp_labels,
                                           return_dict, enrollments)

    def _forward_inference(self, input_features, stno_mask, decoder_input_ids, encoder_outputs, labels, upp_labels,
                           return_dict, enrollments):
        cfg = self.config
        enc_model = self.model.get_encoder()
        if encoder_outputs is None:
            encoder_outputs = enc_model(input_features, stno_mask=stno_mask, enrollments=enrollments)
        enc = encoder_outputs[0] if not isinstance(encoder_outputs, torch.Tensor) else encoder_outputs
        B, T, d = enc.shape
        enc_bf16 = ops.cast_bf16(enc.float())
        hid, hid_bf16 = self.model.decode_teacher_forced(decoder_input_ids, enc_bf16)
        S = decoder_input_ids.shape[1]
        w = self.model.prepare_decoder()
        logits = torch.empty(B, S, cfg.vocab_size, dtype=torch.float32, device=enc.device)
        ops.gemm(hid_bf16, w["proj"], logits.view(B * S, cfg.vocab_size), epilogue=ops.EPI_BIAS_F32)
        loss = None
        if labels is not None:
            labels = labels.to(enc.device)
            flat = logits.view(B * S, cfg.vocab_size)
            if self.soft_label_creator is not None:
                dec_loss = self.soft_label_creator.compute_loss(flat, labels, upp_labels)
            else:  # hard-label fallback, mean over ALL positions (modeling_dicow.py:312-323)
                dec_loss = ops.softlabel_ce(flat, labels, upp_labels.to(enc.device) if upp_labels is not None else None,
                                            soft_mode=False)
            if cfg.ctc_weight > 0.0:
                enc_logits = enc_model.ctc_logits_from_hidden(enc_bf16, B, T)
                ctc = enc_model.get_loss(enc_logits, self._ctc_labels(labels))
                loss = (1 - cfg.ctc_weight) * dec_loss + cfg.ctc_weight * ctc
            else:
                loss = dec_loss
        if return_dict is False:
            out = (logits, enc)
            return ((loss,) + out) if loss is not None else out
        return Seq2SeqLMOutput(loss=loss, logits=logits, encoder_last_hidden_state=enc)

    # ---- greedy decoding of one 30 s window batch ----------------------------------------------------------------
    def _decode_step(self, st: _GreedyState, w: dict, sample: bool, gen: dict) -> None:
        """One token for every row: fixed launch sequence, position read from the device scalar ``st.pos``.
        Per decoder layer (HF:modeling_whisper.py:449-506): LayerNorm, q | k,v as ONE projection whose k,v columns are
        appended to the cache, attention over the cache, out_proj + residual, ... -- the linear layers are
        ops.decode_linear (weights requested first, A staged once in shared memory, K split over a cluster for the
        small-N layers)."""
        cfg = self.config
        d, H, B, T = cfg.d_model, cfg.decoder_attention_heads, st.B, st.T
        if not self.fused_decode_step or d % 32:
            if st.beams > 1:
                raise NotImplementedError("beam search runs on the fused decode step (d_model % 32 == 0)")
            return self._decode_step_unfused(st, w, sample, gen)
        # LayerNorm: the few-rows LayerNorm kernel in front of the linear kernel (default), or ("ln_prologue") as the
        # prologue of the linear kernel, every CTA normalising the <= 32 rows itself (a row is held in registers: d <= 1280).
        # The prologue is measured slower in both of its forms -- shared over a cluster of 8 through DSMEM (round 1: 15-17 us
        # per layer linear) and per CTA (round 2: 10.9-14.4 us against 2.1 + 4.0-6.3 us) -- ~160 CTAs re-reading the same
        # 80 KB of fp32 rows from L2 cost more than the LayerNorm launch they replace
        ln_prologue = self.fused_decode_step == "ln_prologue" and d <= 1280 and B <= 32

        def ln_linear(W, out, g, b, **kw):
            if ln_prologue:
                return ops.decode_linear(W, out, x=st.x, gamma=g, beta=b, **kw)
            ops.fddt_layernorm(st.x, gamma=g, beta=b, ln_out_bf16=st.ln)
            return ops.decode_linear(W, out, A=st.ln, **kw)

        if self._megakernel_ok(st):
            self._decode_layers_mega(st, w)
            if sample:
                ln_linear(w["proj"], st.logits, w["lnf_g"], w["lnf_b"], epilogue=ops.EPI_BIAS_F32)
                self._select_token(st, gen)
            ops.advance(st.pos, 1)
            return
        ops.embed_tokens(st.ids, w["tok"], w["pos"], st.x, S=1, pos=st.pos)
        for li, e in enumerate(w["layers"]):
            s, c = e["self"], e["cross"]
            kvc = st.self_kv[li]
            ln_linear(s["wqkv"], st.q, e["ln1_g"], e["ln1_b"], epilogue=ops.EPI_BIAS_BF16, bias=s["bqkv"], out2=kvc,
                      n_split=d, ldo2=st.S_max * 2 * d, pos=st.pos, pos_stride=2 * d)
            ops.decode_attention(st.q, kvc, kvc[:, :, d:], st.ctx, B=B, H=H, Tk=0, kv_row_stride=2 * d,
                                 kv_batch_stride=st.S_max * 2 * d, pos=st.pos, ancestry=st.ancestry)
            ops.decode_linear(s["wo"], st.x, A=st.ctx, epilogue=ops.EPI_RESIDUAL_F32, bias=s["bo"], resid=st.x)
            ln_linear(c["wq"], st.q, e["ln2_g"], e["ln2_b"], epilogue=ops.EPI_BIAS_BF16, bias=c["bq"])
            ckv = st.cross_kv[li]
            ops.decode_attention(st.q, ckv, ckv[..., 64:], st.ctx, B=B, H=H, Tk=T, kv_row_stride=128,
                                 kv_batch_stride=H * T * 128, kv_head_stride=T * 128, kv_batch_div=st.beams)
            ops.decode_linear(c["wo"], st.x, A=st.ctx, epilogue=ops.EPI_RESIDUAL_F32, bias=c["bo"], resid=st.x)
            ln_linear(e["w1"], st.h, e["ln3_g"], e["ln3_b"], epilogue=ops.EPI_BIAS_GELU_BF16, bias=e["b1"])
            ops.decode_linear(e["w2"], st.x, A=st.h, epilogue=ops.EPI_RESIDUAL_F32, bias=e["b2"], resid=st.x)
        if sample:
            ln_linear(w["proj"], st.logits, w["lnf_g"], w["lnf_b"], epilogue=ops.EPI_BIAS_F32)
            self._select_token(st, gen)
        ops.advance(st.pos, 1)

    def _megakernel_ok(self, st: _GreedyState) -> bool:
        cfg = self.config
        d, ffn = cfg.d_model, cfg.decoder_ffn_dim
        return bool(self.decode_megakernel) and st.beams == 1 and st.B <= 16 and d % 64 == 0 and d <= 1280 and \
            cfg.decoder_attention_heads * 64 == d and ffn % 32 == 0 and ffn <= 5120

    def _decode_layers_mega(self, st: _GreedyState, w: dict) -> None:
        """embedding + every decoder layer of the step in one persistent kernel (csrc/decode_mega.cu)"""
        cfg = self.config
        if st.mega_table is None or st.mega_weights is not w:
            rows = []
            for li, e in enumerate(w["layers"]):
                s, c = e["self"], e["cross"]
                rows.append({"ln1_g": e["ln1_g"], "ln1_b": e["ln1_b"], "ln2_g": e["ln2_g"], "ln2_b": e["ln2_b"],
                             "ln3_g": e["ln3_g"], "ln3_b": e["ln3_b"], "wqkv": s["wqkv"], "bqkv": s["bqkv"],
                             "wo_self": s["wo"], "bo_self": s["bo"], "wq_cross": c["wq"], "bq_cross": c["bq"],
                             "wo_cross": c["wo"], "bo_cross": c["bo"], "w1": e["w1"], "b1": e["b1"], "w2": e["w2"], "b2": e["b2"],
                             "self_kv": st.self_kv[li], "cross_kv": st.cross_kv[li]})
            st.mega_table = ops.decode_layer_table(rows, st.x.device)
            st.mega_weights = w
        ops.decode_layers(st.mega_table, B=st.B, d=cfg.d_model, H=cfg.decoder_attention_heads, ffn=cfg.decoder_ffn_dim,
                          L=len(w["layers"]), T=st.T, S_max=st.S_max, vocab=cfg.vocab_size, ids=st.ids, tok=w["tok"],
                          posw=w["pos"], pos=st.pos, x=st.x, q=st.q, ctx=st.ctx, hidden=st.h, barrier=st.mega_bar,
                          workspace=st.mega_ws)

    def _select_token(self, st: _GreedyState, gen: dict) -> None:
        """logits -> next token of every row.  Attention-only greedy: the fused rules + argmax kernel.  Joint CTC /
        attention (generation_config.ctc_weight > 0; generation.py:250-268): the rules kernel only materialises the
        processed scores, the CTC step log-softmaxes them, scores the top-k candidates' prefixes and selects."""
        ctc = gen.get("ctc")
        rules = {k: v for k, v in gen.items() if k not in ("ctc", "beam")}
        beam = gen.get("beam")
        if beam is not None:
            return self._beam_select(st, rules, ctc, beam)
        if ctc is None:
            ops.logits_rules_argmax(st.logits, st.ids, st.unfinished, pos=st.pos, **rules)
            return
        ops.logits_rules_argmax(st.logits, st.ids, st.unfinished, pos=st.pos, processed_scores=st.proc, no_select=True,
                                **rules)
        ops.ctc_joint_step(st.ctc, st.proc, st.ids, st.unfinished, pos=st.pos, bos=ctc["bos"], eos=rules["eos"],
                           pad=rules["pad"], first_timestamp=rules["ts_begin"], prefix_len=ctc["prefix_len"],
                           ctc_weight=ctc["weight"])

    def _beam_select(self, st: _GreedyState, rules: dict, ctc: Optional[dict], beam: dict) -> None:
        """one beam-search step after the logits (generation.py:1003-1088): processed scores of every hypothesis, their
        log-softmax normaliser (over the RAW logits: beam search normalises before the processors) and top-k candidates,
        CTC prefix scores of the candidates (ctc_weight > 0), then selection across the beams of each utterance,
        finished-set bookkeeping and the re-linking of sequences / ancestry / CTC states -- all on the device."""
        w = float(ctc["weight"]) if ctc is not None else 0.0
        joint = st.ctc if ctc is not None else st.cand
        ops.logits_rules_argmax(st.logits, st.ids, st.unfinished, pos=st.pos, processed_scores=st.proc, no_select=True,
                                **rules)
        ops.ctc_joint_step(joint, st.proc, st.ids, st.unfinished, pos=st.pos, bos=beam["bos"], eos=rules["eos"],
                           pad=rules["pad"], first_timestamp=rules["ts_begin"], prefix_len=beam["prefix_len"], ctc_weight=w,
                           raw_logits=st.logits, score_only=True)
        ops.beam_step(U=st.B // st.beams, NB=st.beams, processed_scores=st.proc, joint=joint, ctc_weight=w,
                      run_score=st.run_score, fin_score=st.fin_score, fin_flag=st.fin_flag, unsat=st.unsat, ids=st.ids,
                      fin_ids=st.fin_ids, ids_tmp=st.ids_tmp, ancestry=st.ancestry, ancestry_tmp=st.ancestry_tmp, pos=st.pos,
                      eos=rules["eos"], pad=rules["pad"], first_timestamp=rules["ts_begin"], max_length=beam["max_length"],
                      prompt_len=beam["prompt_len"], length_penalty=beam["length_penalty"],
                      early_stopping=beam["early_stopping"], scratch_i32=st.scratch_i32, scratch_f32=st.scratch_f32,
                      flags=st.flags, ctc_r_tmp=st.ctc_r_tmp)

    @torch.no_grad()
    def beam_decode_window(self, enc_hidden: torch.Tensor, prompt: torch.Tensor, max_total_len: int, gen: dict, *,
                           num_beams: int, length_penalty: float = 1.0, early_stopping=False, ctc: Optional[dict] = None,
                           top_k: int = 500) -> torch.Tensor:
        """Beam search over one batch of 30 s windows (DiCoWGenerationMixin._beam_search, generation.py:815-1154), with the
        joint CTC / attention rescoring of the published recipe when ``ctc`` is given (configs/decode/*_beam_joint.yaml:
        5 beams, ctc_weight 0.2, length_penalty 0.1).  Rows are utterance-major hypotheses; the beams of an utterance share
        its cross-attention K/V; the self-attention cache is never re-ordered (ancestry table).  Returns the best
        finished sequence per utterance, int64 [U, n], padded with ``pad``."""
        cfg = self.config
        dev = enc_hidden.device
        U, T, d = enc_hidden.shape
        NB = int(num_beams)
        B = U * NB
        P = prompt.shape[1]
        max_total_len = min(max_total_len, cfg.max_target_positions)
        if B > 64:  # the step kernels stage at most 64 hypothesis rows: larger batches run in chunks of floor(64 / beams) windows
            per = max(1, 64 // NB)
            outs = []
            for lo in range(0, U, per):
                sl = slice(lo, min(U, lo + per))
                sub = None if ctc is None else dict(ctc, logits=ctc["logits"][sl])
                outs.append(self.beam_decode_window(enc_hidden[sl], prompt[sl], max_total_len, gen, num_beams=num_beams,
                                                    length_penalty=length_penalty, early_stopping=early_stopping, ctc=sub,
                                                    top_k=top_k))
            n = max(o.shape[1] for o in outs)
            return torch.cat([torch.nn.functional.pad(o, (0, n - o.shape[1]), value=int(gen["pad"])) for o in outs], dim=0)
        w = self.model.prepare_decoder()
        key = (dev.index, B, T, NB)
        st = self._greedy.get(key, lambda: _GreedyState(self, B, T, dev, beams=NB))
        enc_bf16 = enc_hidden if enc_hidden.dtype == torch.bfloat16 else ops.cast_bf16(enc_hidden.float())
        encf = enc_bf16.reshape(U * T, d)
        H = cfg.decoder_attention_heads
        for li, e in enumerate(w["layers"]):
            ops.gemm(encf, e["cross"]["wkv"], st.cross_kv_rows, epilogue=ops.EPI_BIAS_BF16, bias=e["cross"]["bkv"])
            ops.kv_to_head_major(st.cross_kv_rows, st.cross_kv[li], B=U, T=T, H=H)
        if st.weights is not w:
            st.graphs.clear()
            st.weights = w
        # ---- state of a new window (generation.py:940-975) ----
        st.ids.zero_()
        st.ids[:, :P] = prompt.to(device=dev, dtype=torch.int64).repeat_interleave(NB, dim=0)
        st.fin_ids.fill_(gen["pad"])
        st.fin_ids[:, :P] = st.ids[:, :P]
        st.pos.zero_()
        st.unfinished.fill_(1)
        st.run_score.fill_(-1.0e9)
        st.run_score.view(U, NB)[:, 0] = 0.0
        st.fin_score.fill_(-1.0e9)
        st.fin_flag.zero_()
        st.unsat.fill_(1)
        st.flags.zero_()
        st.ancestry.copy_(torch.arange(B, dtype=torch.int32, device=dev)[:, None].expand(B, st.S_max))
        if st.proc is None:
            st.proc = torch.empty(B, cfg.vocab_size, dtype=torch.float32, device=dev)
        k_eff = min(int(top_k), gen["ts_begin"])
        gen = dict(gen, begin_index=P)
        if ctc is not None:
            lg = ctc["logits"].float().repeat_interleave(NB, dim=0).contiguous()  # generation.py:253
            ckey = (tuple(lg.shape), k_eff, tuple(sorted((ctc.get("upper_cased") or {}).items())))
            if st.ctc is None or st.ctc_key != ckey:
                st.ctc = ops.CtcJointState(lg, top_k=k_eff, upper_cased=ctc.get("upper_cased"))
                st.ctc_key = ckey
                st.ctc_r_tmp = torch.empty_like(st.ctc.r_prev)
                st.ctc_gen += 1
                st.graphs.clear()
            else:
                st.ctc.reset(lg)
            gen["ctc"] = {"weight": float(ctc["weight"]), "state": st.ctc_gen}
        elif st.cand is None or st.cand.K != k_eff:
            st.cand = ops.CandidateState(B, k_eff, dev)
            st.graphs.clear()
        gen["beam"] = {"bos": int(cfg.decoder_start_token_id), "prefix_len": int(ctc["prefix_len"]) if ctc is not None else P,
                       "max_length": int(max_total_len), "prompt_len": P, "length_penalty": float(length_penalty),
                       "early_stopping": early_stopping}

        def run(sample: bool):
            if not self.use_cuda_graphs:
                self._decode_step(st, w, sample, gen)
                return
            gkey = (sample, P, self.fused_decode_step,
                    tuple(sorted((k, v.data_ptr() if isinstance(v, torch.Tensor) else
                                  (tuple(sorted(v.items())) if isinstance(v, dict) else v)) for k, v in gen.items())))
            g = st.graphs.get(gkey)
            if g is None:
                # warm-up outside capture, then restore everything the step changed
                keep = [t.clone() for t in self._beam_mutable(st)]
                self._decode_step(st, w, sample, gen)
                torch.cuda.synchronize(dev)
                for t, k0 in zip(self._beam_mutable(st), keep):
                    t.copy_(k0)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._decode_step(st, w, sample, gen)
                st.graphs[gkey] = g
                for t, k0 in zip(self._beam_mutable(st), keep):
                    t.copy_(k0)
            g.replay()

        for _ in range(P - 1):
            run(False)
        early_true = early_stopping is True
        for _ in range(max_total_len - P):
            run(True)
            f = st.flags.cpu()  # the loop condition is a property of the whole batch (generation.py:1101-1106)
            improvement = bool(f[:, 2].any())
            open_beam = not (bool(f[:, 1].all()) and early_true)
            valid = bool(f[:, 0].any())
            if not (improvement and open_beam and valid):
                break
        best = st.fin_ids.view(U, NB, -1)[:, 0, :max_total_len].clone()
        return best

    @staticmethod
    def _beam_mutable(st: _GreedyState):
        ts = [st.ids, st.pos, st.unfinished, st.run_score, st.fin_score, st.fin_flag, st.unsat, st.fin_ids, st.ancestry,
              st.flags]  # (the warm-up step's K/V rows are rewritten by the captured step at the same position)
        if st.ctc is not None:
            ts += [st.ctc.r_prev, st.ctc.score_prev]
        return ts

    def _decode_step_unfused(self, st: _GreedyState, w: dict, sample: bool, gen: dict) -> None:
        """the same step with one kernel per operation (LayerNorm, q, k|v, ... 13 per layer): kept as the comparison
        baseline of the fused step (tests/test_gpu_decoder.py, tools/bench_decode.py --unfused)"""
        cfg = self.config
        d, H, B, T = cfg.d_model, cfg.decoder_attention_heads, st.B, st.T
        ops.embed_tokens(st.ids, w["tok"], w["pos"], st.x, S=1, pos=st.pos)
        for li, e in enumerate(w["layers"]):
            s, c = e["self"], e["cross"]
            kvc = st.self_kv[li]
            ops.fddt_layernorm(st.x, gamma=e["ln1_g"], beta=e["ln1_b"], ln_out_bf16=st.ln)
            ops.gemm_skinny(st.ln, s["wq"], st.q, epilogue=ops.EPI_BIAS_BF16, bias=s["bq"])
            ops.gemm_skinny(st.ln, s["wkv"], kvc, epilogue=ops.EPI_BIAS_BF16, bias=s["bkv"], ldo=st.S_max * 2 * d,
                            pos=st.pos, pos_stride=2 * d)
            ops.decode_attention(st.q, kvc, kvc[:, :, d:], st.ctx, B=B, H=H, Tk=0, kv_row_stride=2 * d,
                                 kv_batch_stride=st.S_max * 2 * d, pos=st.pos)
            ops.gemm_skinny(st.ctx, s["wo"], st.x, epilogue=ops.EPI_RESIDUAL_F32, bias=s["bo"], resid=st.x)
            ops.fddt_layernorm(st.x, gamma=e["ln2_g"], beta=e["ln2_b"], ln_out_bf16=st.ln)
            ops.gemm_skinny(st.ln, c["wq"], st.q, epilogue=ops.EPI_BIAS_BF16, bias=c["bq"])
            ckv = st.cross_kv[li]
            ops.decode_attention(st.q, ckv, ckv[..., 64:], st.ctx, B=B, H=H, Tk=T, kv_row_stride=128,
                                 kv_batch_stride=H * T * 128, kv_head_stride=T * 128)
            ops.gemm_skinny(st.ctx, c["wo"], st.x, epilogue=ops.EPI_RESIDUAL_F32, bias=c["bo"], resid=st.x)
            ops.fddt_layernorm(st.x, gamma=e["ln3_g"], beta=e["ln3_b"], ln_out_bf16=st.ln)
            ops.gemm_skinny(st.ln, e["w1"], st.h, epilogue=ops.EPI_BIAS_GELU_BF16, bias=e["b1"])
            ops.gemm_skinny(st.h, e["w2"], st.x, epilogue=ops.EPI_RESIDUAL_F32, bias=e["b2"], resid=st.x)
        if sample:
            ops.fddt_layernorm(st.x, gamma=w["lnf_g"], beta=w["lnf_b"], ln_out_bf16=st.ln)
            ops.gemm_skinny(st.ln, w["proj"], st.logits, epilogue=ops.EPI_BIAS_F32)
            self._select_token(st, gen)
        ops.advance(st.pos, 1)

    @torch.no_grad()
    def greedy_decode_window(self, enc_hidden: torch.Tensor, prompt: torch.Tensor, max_total_len: int, gen: dict,
                             return_first_logits: bool = False, ctc: Optional[dict] = None):
        """Greedy branch of DiCoWGenerationMixin._sample (generation.py:707-782) for one batch of 30 s windows.
        enc_hidden [B, T, d] (fp32 or bf16), prompt int64 [B, P] (forced init tokens).  Returns int64 ids [B, n] on
        the device (prompt + generated, finished rows padded).  ``ctc`` = {"logits": fp32 [B, T', V + 1] CTC logits of the
        window, "weight", "prefix_len", "bos", "top_k", "upper_cased"} switches to joint CTC / attention selection."""
        cfg = self.config
        dev = enc_hidden.device
        B, T, d = enc_hidden.shape
        P = prompt.shape[1]
        max_total_len = min(max_total_len, cfg.max_target_positions)
        w = self.model.prepare_decoder()
        if B > 64:  # the step kernels stage at most 64 rows: larger batches decode in chunks of 64 windows
            outs, firsts = [], []
            for lo in range(0, B, 64):
                sl = slice(lo, min(B, lo + 64))
                sub = None if ctc is None else dict(ctc, logits=ctc["logits"][sl])
                r = self.greedy_decode_window(enc_hidden[sl], prompt[sl], max_total_len, gen, return_first_logits, sub)
                outs.append(r[0] if return_first_logits else r)
                if return_first_logits:
                    firsts.append(r[1])
            n = max(o.shape[1] for o in outs)
            ids = torch.cat([torch.nn.functional.pad(o, (0, n - o.shape[1]), value=int(gen["pad"])) for o in outs], dim=0)
            return (ids, torch.cat(firsts, dim=0)) if return_first_logits else ids
        # rows of the decode state: B rounded up to a bucket (padding rows start finished and decode nothing that is read)
        Bs = B if ctc is not None else _bucket_rows(B)
        key = (dev.index, Bs, T)
        st = self._greedy.get(key, lambda: _GreedyState(self, Bs, T, dev))
        enc_bf16 = enc_hidden if enc_hidden.dtype == torch.bfloat16 else ops.cast_bf16(enc_hidden.float())
        encf = enc_bf16.reshape(B * T, d)
        H = cfg.decoder_attention_heads
        for li, e in enumerate(w["layers"]):  # cross-attention K/V once per window (HF caches them after step 0)
            ops.gemm(encf, e["cross"]["wkv"], st.cross_kv_rows[:B * T], epilogue=ops.EPI_BIAS_BF16, bias=e["cross"]["bkv"])
            ops.kv_to_head_major(st.cross_kv_rows, st.cross_kv[li], B=B, T=T, H=H)
        if st.weights is not w:  # parameters changed since capture: the graphs hold stale weight pointers
            st.graphs.clear()
            st.weights = w
        st.ids.zero_()
        st.ids[:B, :P] = prompt.to(device=dev, dtype=torch.int64)
        st.pos.zero_()
        st.unfinished.fill_(1)
        if Bs > B:
            st.ids[B:, :P] = st.ids[:1, :P]
            st.unfinished[B:] = 0
        gen = dict(gen, begin_index=P)
        if ctc is not None:
            lg = ctc["logits"].float().contiguous()
            ckey = (tuple(lg.shape), int(ctc.get("top_k", 500)), tuple(sorted((ctc.get("upper_cased") or {}).items())))
            if st.ctc is None or st.ctc_key != ckey:  # buffers are kept across windows: captured graphs hold their pointers
                st.ctc = ops.CtcJointState(lg, top_k=ckey[1], upper_cased=ctc.get("upper_cased"))
                st.ctc_key = ckey
                st.proc = torch.empty(B, cfg.vocab_size, dtype=torch.float32, device=dev)
                st.ctc_gen += 1
                st.graphs.clear()
            else:
                st.ctc.reset(lg)
            gen["ctc"] = {"bos": int(ctc["bos"]), "prefix_len": int(ctc["prefix_len"]), "weight": float(ctc["weight"]),
                          "state": st.ctc_gen}

        def run(sample: bool):
            if not self.use_cuda_graphs:
                self._decode_step(st, w, sample, gen)
                return
            gkey = (sample, P, self.fused_decode_step,
                    tuple(sorted((k, v.data_ptr() if isinstance(v, torch.Tensor) else
                                  (tuple(sorted(v.items())) if isinstance(v, dict) else v)) for k, v in gen.items())))
            g = st.graphs.get(gkey)
            if g is None:
                # warm-up launch outside capture (kernel attributes, lazy module load), then restore the state it touched
                ids0, pos0, unf0 = st.ids.clone(), st.pos.clone(), st.unfinished.clone()
                self._decode_step(st, w, sample, gen)
                torch.cuda.synchronize(dev)
                st.ids.copy_(ids0), st.pos.copy_(pos0), st.unfinished.copy_(unf0)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._decode_step(st, w, sample, gen)
                st.graphs[gkey] = g
                st.ids.copy_(ids0), st.pos.copy_(pos0), st.unfinished.copy_(unf0)
            g.replay()

        for _ in range(P - 1):  # prompt tokens: fill the self-attention cache only
            run(False)
        first_logits = None
        n_new = max_total_len - P
        for step in range(n_new):
            run(True)
            if step == 0 and return_first_logits:
                first_logits = st.logits[:B].clone()
            if (step & 7) == 7 and not bool(st.unfinished.any().item()):  # host check every 8 tokens only
                n_new = step + 1
                break
        ids = st.ids[:B, :P + n_new].clone()
        # rows that finished early were padded by the kernel; trim columns that are padding for every row
        return (ids, first_logits) if return_first_logits else ids

    def _suppress_bitmap(self, tokens, dev) -> Optional[torch.Tensor]:
        """device bitmap of SuppressTokensLogitsProcessor's ids, cached so captured graphs keep a valid pointer"""
        if not tokens:
            return None
        key = (dev.index, tuple(int(t) for t in tokens))
        cache = self.__dict__.setdefault("_suppress_cache", {})
        if key not in cache:
            cache[key] = ops.suppress_bitmap(key[1], self.config.vocab_size, dev)
        return cache[key]

    # ---- generate() -----------------------------------------------------------------------------------------
    def _generation_settings(self, generation_config, kwargs) -> dict:
        gc = generation_config if generation_config is not None else self.generation_config
        def get(name, default=None):
            if name in kwargs and kwargs[name] is not None:
                return kwargs[name]
            v = getattr(gc, name, None)
            return default if v is None else v
        num_beams = int(get("num_beams", 1) or 1)
        if num_beams > 8:
            raise NotImplementedError("beam search with more than 8 beams")
        if int(get("num_return_sequences", 1) or 1) != 1:
            raise NotImplementedError("num_return_sequences > 1")
        ctc_weight = float(get("ctc_weight", 0) or 0)
        # logits processors the reference switches off (update_generation_config, src/utils/general.py:19-37:
        # begin_suppress_tokens=None, repetition_penalty from the decoding args, default None) are not built: refuse
        # instead of silently decoding without them
        if get("begin_suppress_tokens"):
            raise NotImplementedError("begin_suppress_tokens is not applied by the B200 decode step; the reference sets it "
                                      "to None (src/utils/general.py:26) -- do the same on this generation config")
        if get("repetition_penalty") not in (None, 1.0) or get("no_repeat_ngram_size") not in (None, 0):
            raise NotImplementedError("repetition_penalty / no_repeat_ngram_size are not applied by the B200 decode step")
        # explicit requests for HF long-form features the reference's recipes never use (temperature fallback with its
        # thresholds, prompts, token-level timestamps): refuse rather than return something else than what was asked for
        temp = kwargs.get("temperature")
        if isinstance(temp, (list, tuple)) and len(temp) > 1:
            raise NotImplementedError("temperature fallback (a tuple of temperatures) is not built")
        for name in ("no_speech_threshold", "logprob_threshold", "compression_ratio_threshold", "prompt_ids"):
            if kwargs.get(name) is not None:
                raise NotImplementedError(f"{name} is not supported by the B200 generate()")
        if kwargs.get("return_token_timestamps"):
            raise NotImplementedError("return_token_timestamps needs cross-attention weights, which the fused path does not produce")
        if get("do_sample", False):
            raise ValueError("Provided generation mode is not supported (greedy only)")
        ts_begin = get("no_timestamps_token_id")
        if ts_begin is None:
            raise ValueError("generation_config.no_timestamps_token_id is required (Whisper generation config)")
        eos = get("eos_token_id", self.config.eos_token_id)
        eos = eos[0] if isinstance(eos, (list, tuple)) else eos
        pad = get("pad_token_id", self.config.pad_token_id)
        return {"eos": int(eos), "pad": int(pad if pad is not None else eos), "no_timestamps": int(ts_begin),
                "ts_begin": int(ts_begin) + 1, "suppress_tokens": get("suppress_tokens"),
                "return_timestamps": bool(get("return_timestamps", True)),
                "max_initial_timestamp_index": get("max_initial_timestamp_index"),
                "max_new_tokens": get("max_new_tokens"), "max_length": get("max_length", self.config.max_target_positions),
                "forced_decoder_ids": get("forced_decoder_ids"), "ctc_weight": ctc_weight, "num_beams": num_beams,
                "length_penalty": float(get("length_penalty", 1.0)), "early_stopping": get("early_stopping", False),
                "ctc_tokens_to_score": int(get("ctc_tokens_to_score", 500) or 500)}

    @torch.no_grad()
    def generate(self, input_features: Optional[torch.Tensor] = None, generation_config=None,
                 condition_on_prev_tokens: Optional[bool] = None, assistant_model=None,
                 attention_mask: Optional[torch.Tensor] = None, stno_mask: Optional[torch.Tensor] = None,
                 forced_decoder_ids=None, enrollments: Optional[dict] = None, return_segments: bool = False,
                 **kwargs):
        """src/models/dicow/generation.py:536-564 over HF WhisperGenerationMixin.generate (long-form seek loop)."""
        if condition_on_prev_tokens:
            raise NotImplementedError("Current version does not support conditioning")  # generation.py:543-544
        if assistant_model is not None:
            raise NotImplementedError("assisted generation is not supported")
        cfg = self.config
        if forced_decoder_ids is not None:
            kwargs["forced_decoder_ids"] = forced_decoder_ids
        gs = self._generation_settings(generation_config, kwargs)
        fdi = gs["forced_decoder_ids"]
        if fdi is None:
            fdi = getattr(cfg, "forced_decoder_ids", None)
        dev = input_features.device
        if fdi is None:  # generation.py:145-147 -> HF _retrieve_init_tokens: <|sot|> <|lang|> <|task|> [<|notimestamps|>]
            self.stno_mask = stno_mask
            fdi = self._init_tokens_without_forced_ids(input_features, stno_mask, enrollments, generation_config, kwargs, gs)
        init_tokens = torch.as_tensor(fdi, dtype=torch.int64, device=dev)
        if init_tokens.dim() == 1:
            init_tokens = init_tokens[None].expand(input_features.shape[0], -1)
        B0 = input_features.shape[0]
        P = init_tokens.shape[1]
        if gs["max_new_tokens"] is not None:
            if gs["max_new_tokens"] + P > cfg.max_target_positions:
                raise ValueError(f"The length of `decoder_input_ids`, including special start tokens, prompt tokens, and "
                                 f"previous tokens, is {P},  and `max_new_tokens` is {gs['max_new_tokens']}. Thus, the "
                                 f"combined length of `decoder_input_ids` and `max_new_tokens` is: "
                                 f"{gs['max_new_tokens'] + P}. This exceeds the `max_target_positions` of the Whisper "
                                 f"model: {cfg.max_target_positions}.")
            max_total = P + int(gs["max_new_tokens"])
        else:
            max_total = int(gs["max_length"])
        self.stno_mask = stno_mask
        enc = self.model.get_encoder()
        input_stride = 2
        num_segment_frames = input_stride * cfg.max_source_positions
        time_precision = 0.02
        total_frames = input_features.shape[-1]
        if attention_mask is not None:
            max_frames = attention_mask.sum(-1).cpu().to(torch.long)
        else:
            max_frames = torch.full((B0,), total_frames, dtype=torch.long)
        seek = torch.zeros(B0, dtype=torch.long)
        rules = dict(eos=gs["eos"], pad=gs["pad"], no_timestamps=gs["no_timestamps"], ts_begin=gs["ts_begin"],
                     max_initial_timestamp_index=gs["max_initial_timestamp_index"],
                     timestamp_rules=gs["return_timestamps"],
                     suppress_bitmap=self._suppress_bitmap(gs["suppress_tokens"], dev))
        segments: List[List[dict]] = [[] for _ in range(B0)]
        batch_idx_map = list(range(B0))
        feats = input_features
        enr_cache, enr_rows = None, {}
        spec = None
        speculate = bool(getattr(self, "speculate_next_window", False)) and dev.type == "cuda"
        self.speculation_stats = {"hits": 0, "misses": 0}
        timing: dict = {}  # the speculative pass' SM budget and whether the last pass kept the loop waiting
        while bool((seek < max_frames).any()):
            # drop finished recordings from the batch (HF:_maybe_reduce_batch)
            keep = [i for i, prev in enumerate(batch_idx_map) if seek[prev] < max_frames[prev]]
            if len(keep) != len(batch_idx_map):
                feats = feats[keep]
                batch_idx_map = [batch_idx_map[i] for i in keep]
            cur = len(batch_idx_map)
            time_offset = seek.to(torch.float64) * time_precision / input_stride
            seek_num_frames = (max_frames - seek).clamp(max=num_segment_frames)

            def window_inputs(rows, at):
                """mel / STNO windows of the recordings ``rows`` (positions in ``feats`` and original indices) starting at
                mel frame ``at[prev]``"""
                w_in, w_stno = [], []
                for i, prev in rows:
                    s0 = int(at[prev])
                    n = int(min(int(max_frames[prev]) - s0, num_segment_frames))
                    f = feats[i:i + 1, :, s0:s0 + n]
                    if f.shape[-1] < num_segment_frames:  # HF:_get_input_segment pads the mel with zeros
                        f = torch.nn.functional.pad(f, (0, num_segment_frames - f.shape[-1]))
                    w_in.append(f)
                    if stno_mask is not None:  # generation.py:73-118: STNO index = mel frame // 2, silence-padded
                        v0 = s0 // 2
                        nv = int((max_frames[prev] // 2 - v0).clamp(max=num_segment_frames // 2))
                        m = stno_mask[prev:prev + 1, :, v0:v0 + nv]
                        if m.shape[-1] < num_segment_frames // 2:
                            orig = m.shape[-1]
                            m = torch.nn.functional.pad(m, (0, num_segment_frames // 2 - orig))
                            m[0, 0, orig:] = 1.0
                        w_stno.append(m)
                return torch.cat(w_in, 0), (torch.cat(w_stno, 0) if w_stno else None)

            def enrollment_args(prevs, first_pass):
                """(enrollments, enrollment_kv, capture list) for an encoder pass over the recordings ``prevs``"""
                if not (cfg.use_enrollments and enrollments is not None):
                    return None, None, None
                if enr_cache is not None:
                    # later windows of a recording: the enrollment stream's keys / values of every speaker communication
                    # block are those of its first window (the stream never reads the target stream) -- reuse them
                    rows_of = torch.as_tensor([enr_rows[prev] for prev in prevs], device=dev)
                    return None, [c.index_select(0, rows_of) for c in enr_cache], None
                idx = torch.as_tensor(prevs, device=dev)
                e = {k: v[idx] for k, v in enrollments.items()}
                cap = None
                if first_pass and getattr(self, "cache_enrollment_kv", True) and bool((max_frames > num_segment_frames).any()):
                    cap = []
                return e, None, cap

            seg_in, seg_stno = window_inputs(list(enumerate(batch_idx_map)), seek)
            self.stno_mask_seek = seg_stno
            hidden = None
            if spec is not None:
                # the window encoded ahead (below) is this iteration's window for every recording still in the batch
                if all(prev in spec["rows"] and spec["at"][prev] == int(seek[prev]) for prev in batch_idx_map):
                    timing["waited"] = not spec["done"].query()  # the decode finished first: the pass wants more SMs
                    torch.cuda.current_stream(dev).wait_event(spec["done"])
                    pick = [spec["rows"][prev] for prev in batch_idx_map]
                    hidden = spec["hidden"] if pick == list(range(spec["hidden"].shape[0])) else \
                        spec["hidden"].index_select(0, torch.as_tensor(pick, device=dev))
                    hidden.record_stream(torch.cuda.current_stream(dev))
                    self.speculation_stats["hits"] += 1
                else:
                    self.speculation_stats["misses"] += 1
                spec = None
            if hidden is None:
                enr, enr_kv, capture = enrollment_args(batch_idx_map, True)
                hidden = enc(seg_in, stno_mask=seg_stno, enrollments=enr, enrollment_kv=enr_kv,
                             capture_enrollment_kv=capture).last_hidden_state
                if capture:
                    enr_cache, enr_rows = capture, {prev: i for i, prev in enumerate(batch_idx_map)}
            if speculate:
                # Encode the windows at seek + 3000 on a second stream, on part of the SMs, while the latency-bound decode steps
                # of this window run: the seek of window n + 1 is only known after window n has been decoded
                # (generation.py:415-534), but it IS seek + 3000 whenever the window closes on a single timestamp, carries no
                # timestamp pair, or timestamps are off -- and the encoder output of a window does not depend on which other
                # windows share its batch, so a hit is bit-identical to encoding after the fact.
                ahead = [(i, prev) for i, prev in enumerate(batch_idx_map)
                         if int(seek[prev]) + num_segment_frames < int(max_frames[prev])]
                if ahead and not (cfg.use_enrollments and enrollments is not None and enr_cache is None
                                  and getattr(self, "cache_enrollment_kv", True)):
                    at = {prev: int(seek[prev]) + num_segment_frames for _, prev in ahead}
                    side = self._speculation_stream(dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    feats.record_stream(side)  # read by the side stream after this iteration may have dropped it
                    with torch.cuda.stream(side):
                        a_in, a_stno = window_inputs(ahead, at)
                        a_enr, a_kv, _ = enrollment_args([prev for _, prev in ahead], False)
                        with ops.sm_budget(dev, self._speculation_budget(dev, len(ahead), max_total - P, timing)):
                            a_hidden = enc(a_in, stno_mask=a_stno, enrollments=a_enr, enrollment_kv=a_kv).last_hidden_state
                        done = torch.cuda.Event()
                        done.record(side)
                    spec = {"rows": {prev: j for j, (_, prev) in enumerate(ahead)}, "at": at, "hidden": a_hidden, "done": done}
            ctc = None
            if gs["ctc_weight"] > 0:  # generation.py:49-51, 250-268: the encoder's CTC posteriors rescore every step
                if not hasattr(enc, "lm_head"):
                    raise ValueError("generation_config.ctc_weight > 0 needs an encoder with a CTC head (config.ctc_weight > 0)")
                Bw, Tw, _ = hidden.shape
                self.encoder_logits = enc.ctc_logits_from_hidden(ops.cast_bf16(hidden.float()), Bw, Tw)
                tok = self.tokenizer
                ctc = {"logits": self.encoder_logits, "weight": gs["ctc_weight"], "top_k": gs["ctc_tokens_to_score"],
                       "bos": cfg.decoder_start_token_id,
                       "prefix_len": len(tok.prefix_tokens) if tok is not None else P,
                       "upper_cased": dict(getattr(tok, "upper_cased_tokens", None) or {}) if tok is not None else None}
            prompts = init_tokens[torch.as_tensor(batch_idx_map, device=dev)]
            if gs["num_beams"] > 1:  # generation.py:815-1154; at most 64 hypotheses per decode batch
                NB = gs["num_beams"]
                per = max(1, 64 // NB)
                parts = []
                for c0 in range(0, hidden.shape[0], per):
                    sub = None if ctc is None else dict(ctc, logits=ctc["logits"][c0:c0 + per])
                    parts.append(self.beam_decode_window(hidden[c0:c0 + per], prompts[c0:c0 + per], max_total, rules,
                                                         num_beams=NB, length_penalty=gs["length_penalty"],
                                                         early_stopping=gs["early_stopping"], ctc=sub,
                                                         top_k=gs["ctc_tokens_to_score"]))
                n = max(q.shape[1] for q in parts)
                ids = torch.cat([torch.nn.functional.pad(q, (0, n - q.shape[1]), value=gs["pad"]) for q in parts], 0)
            else:
                ids = self.greedy_decode_window(hidden, prompts, max_total, rules, ctc=ctc)
            self.stno_mask_seek = None
            ids_host = ids.cpu()
            for i, prev in enumerate(batch_idx_map):
                seq = ids_host[i, P:]
                # strip padding but keep one eos, then drop the eos (HF:generate_with_fallback post-processing)
                if seq.numel() and int(seq[-1]) == gs["pad"]:
                    npad = int((seq == gs["pad"]).sum())
                    if gs["pad"] == gs["eos"]:
                        npad -= 1
                    if npad:
                        seq = seq[:-npad]
                if seq.numel() and int(seq[-1]) == gs["eos"]:
                    seq = seq[:-1]
                segs, offset = self._retrieve_segment(seq, float(time_offset[prev]), gs["ts_begin"],
                                                      int(seek_num_frames[prev]), time_precision, input_stride)
                seek[prev] += offset
                segments[prev] += segs
        # pad the concatenated segment tokens on the right (HF:_pad_to_max_length)
        seqs = [torch.cat([s["tokens"] for s in sl]) if sl else torch.zeros(0, dtype=torch.int64) for sl in segments]
        n = max([int(s.numel()) for s in seqs] + [0])
        out = torch.full((B0, n), gs["pad"], dtype=torch.int64)
        for i, s in enumerate(seqs):
            out[i, :s.numel()] = s
        outputs = {"sequences": out.to(dev), "segments": segments}
        self.encoder_logits = None
        if return_segments:
            return outputs
        if self.tokenizer is not None:
            return self._fix_timestamps_from_segmentation(outputs)
        return outputs["sequences"]

    # ---- long-form speculation -------------------------------------------------------------
    speculate_next_window = False  # generate(): encode the window at seek + 3000 under the decode steps of the current one
    speculation_sms = None         # SMs the speculative encoder pass sizes its persistent kernels for (the rest: decode);
    #                                None: chosen per window from the measured encoder / decode times (_speculation_budget)

    def _speculation_budget(self, dev, windows: int, new_tokens: int, timing: dict) -> int:
        """The speculative pass should end about when the decode steps of the current window do: with too few SMs the next
        iteration waits for it at reduced width, with too many the decode steps starve (measured, 16 recordings x 4 windows:
        64-token windows -- 64 SMs 275 -> 269 ms, 96 SMs 271 -> 247; SE-DiCoW with 128-token windows -- 64 SMs 441 -> 392,
        96 SMs 409).  First window: device x encoder share of (encoder + decode) time, from a FLOP / step-count estimate; every
        later window moves the budget towards the side that finished last (the previous pass was still running when its output
        was needed: more SMs; it had finished: fewer), by 16 SMs at first, half as many after every reversal."""
        if self.speculation_sms is not None:
            return int(self.speculation_sms)
        cfg = self.config
        phys = torch.cuda.get_device_properties(dev).multi_processor_count
        lo, hi = max(2, int(0.22 * phys) & ~1), int(0.8 * phys) & ~1
        b = timing.get("budget")
        if b is None:
            T, d, ffn, L = cfg.max_source_positions, cfg.d_model, cfg.encoder_ffn_dim, cfg.encoder_layers
            enc_ms = L * (2.0 * T * (4 * d * d + 2 * d * ffn) + 4.0 * T * T * d) / 1e12  # ~1 PFLOP/s, per window
            if cfg.use_enrollments and cfg.scb_layers:
                enc_ms *= 1.0 + cfg.scb_layers / max(1, L)
            dec_ms = new_tokens * 0.11 * cfg.decoder_layers / max(1, windows)  # latency-bound steps, shared by the batch
            b = int(phys * (0.1 + enc_ms / max(enc_ms + dec_ms, 1e-6)))
        elif "waited" in timing:
            up = bool(timing.pop("waited"))
            step = timing.get("step", 16)
            if "up" in timing and timing["up"] != up:  # overshot the balance point: smaller steps from here on
                step = max(4, step // 2)
            timing["step"], timing["up"] = step, up
            b += step if up else -step
        b = min(max(b & ~1, lo), hi)
        timing["budget"] = b
        self.speculation_stats.setdefault("sm_budgets", []).append(b)
        return b

    def _speculation_stream(self, dev):
        key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
        st = _SPECULATION_STREAMS.get(key)
        if st is None:
            st = _SPECULATION_STREAMS[key] = torch.cuda.Stream(device=dev)
        return st

    # ---- language detection / prompt construction without forced_decoder_ids ------------------------------------------
    @torch.no_grad()
    def detect_language(self, input_features: Optional[torch.Tensor] = None, encoder_outputs=None, generation_config=None,
                        num_segment_frames: int = 3000, stno_mask: Optional[torch.Tensor] = None,
                        enrollments: Optional[dict] = None) -> torch.Tensor:
        """src/models/dicow/generation.py:151-221: one decoder step on <|startoftranscript|> over the first window, every
        non-language logit masked, argmax.  Returns int64 language token ids [B]."""
        if input_features is None and encoder_outputs is None:
            raise ValueError("You have to specify either `input_features` or `encoder_outputs`")
        if input_features is not None and encoder_outputs is not None:
            raise ValueError("Make sure to specify only one of `input_features` or `encoder_outputs` - not both!")
        gc = generation_config or self.generation_config
        lang_to_id = getattr(gc, "lang_to_id", None)
        if not lang_to_id:
            raise ValueError("detect_language needs generation_config.lang_to_id")
        stno = stno_mask if stno_mask is not None else self.stno_mask
        if input_features is not None:
            B = input_features.shape[0]
            feats = input_features[:, :, :num_segment_frames]
            dev = feats.device
        else:
            enc0 = encoder_outputs[0] if not isinstance(encoder_outputs, torch.Tensor) else encoder_outputs
            B, dev, feats = enc0.shape[0], enc0.device, None
        start = getattr(gc, "decoder_start_token_id", None) or self.config.decoder_start_token_id
        ids = torch.full((B, 1), int(start), dtype=torch.int64, device=dev)
        out = self._forward_inference(feats, stno[:, :, :num_segment_frames // 2] if stno is not None else None, ids,
                                      encoder_outputs, None, None, True, enrollments)
        logits = out.logits[:, -1].float()
        keep = torch.zeros(logits.shape[-1], dtype=torch.bool, device=dev)
        keep[torch.as_tensor(sorted(int(v) for v in lang_to_id.values()), device=dev)] = True
        return logits.masked_fill(~keep, -float("inf")).argmax(-1)

    def _init_tokens_without_forced_ids(self, input_features, stno_mask, enrollments, generation_config, kwargs, gs):
        """HF WhisperGenerationMixin._retrieve_init_tokens (third-party; reached from generation.py:145-147 when no
        forced_decoder_ids are given): <|startoftranscript|>, the language token (given or detected per recording on its
        first window), the task token (default transcribe), <|notimestamps|> unless timestamps are returned."""
        gc = generation_config if generation_config is not None else self.generation_config
        lang_to_id, task_to_id = getattr(gc, "lang_to_id", None), getattr(gc, "task_to_id", None)
        if not lang_to_id or not task_to_id:
            raise ValueError("generate() without forced_decoder_ids needs generation_config.lang_to_id / task_to_id "
                             "(a multilingual Whisper generation config)")
        B = input_features.shape[0]
        start = int(getattr(gc, "decoder_start_token_id", None) or self.config.decoder_start_token_id)
        language = kwargs.get("language", getattr(gc, "language", None))
        task = kwargs.get("task", getattr(gc, "task", None))
        if task is not None and task not in ("translate", "transcribe"):  # HF TASK_IDS
            raise ValueError(f"The `{task}` task is not supported. The task should be one of `['translate', 'transcribe']`")
        if language is not None:
            langs = [language] * B if isinstance(language, str) else list(language)
            lang_ids = []
            for lg in langs:
                tok = lg if lg in lang_to_id else f"<|{lg}|>"
                if tok not in lang_to_id:
                    raise ValueError(f"Unsupported language: {lg}. Language should be one of: {list(lang_to_id)}.")
                lang_ids.append(int(lang_to_id[tok]))
            lang_ids = torch.tensor(lang_ids, dtype=torch.int64)
        else:
            lang_ids = self.detect_language(input_features=input_features, generation_config=gc, stno_mask=stno_mask,
                                            enrollments=enrollments).cpu()
        cols = [torch.full((B,), start, dtype=torch.int64), lang_ids]
        # the task token follows a GIVEN task, or a GIVEN language (default transcribe); with a detected language and no task
        # HF leaves the prompt at <|sot|><|lang|> (generation_whisper.py, "Update init_tokens with task").  The reference's
        # container sets generation_config.task = "transcribe" (src/models/containers.py:59), i.e. the first case.
        if task is not None:
            cols.append(torch.full((B,), int(task_to_id[task]), dtype=torch.int64))
        elif language is not None:
            cols.append(torch.full((B,), int(task_to_id["transcribe"]), dtype=torch.int64))
        if not gs["return_timestamps"]:
            cols.append(torch.full((B,), gs["no_timestamps"], dtype=torch.int64))
        return torch.stack(cols, dim=1)

    @staticmethod
    def _retrieve_segment(seq: torch.Tensor, time_offset: float, timestamp_begin: int, seek_num_frames: int,
                          time_precision: float, input_stride: int):
        """Split one window's tokens on consecutive timestamp pairs and compute the seek advance
        (src/models/dicow/generation.py:415-534).  ``seq`` is a CPU int64 tensor without prompt / eos."""
        toks = seq.tolist()
        is_ts = [t >= timestamp_begin for t in toks]
        single_ts_ending = is_ts[-2:] == [False, True]
        pair_ends = [i + 1 for i in range(len(toks) - 1) if is_ts[i] and is_ts[i + 1]]
        segments = []
        if pair_ends:
            slices = list(pair_ends)
            if single_ts_ending:
                slices.append(len(toks))
            else:
                slices[-1] += 1  # keep the closing timestamp in the last segment
            last = 0
            for k, cur in enumerate(slices):
                part = seq[last:cur]
                is_last = k == len(slices) - 1
                start_pos = int(part[0]) - timestamp_begin
                end_pos = int(part[-1 if (not is_last or single_ts_ending) else -2]) - timestamp_begin
                segments.append({"start": torch.tensor(time_offset + start_pos * time_precision, dtype=torch.float64),
                                 "end": torch.tensor(time_offset + end_pos * time_precision, dtype=torch.float64),
                                 "tokens": part})
                last = cur
            if single_ts_ending:
                offset = seek_num_frames  # no speech after the last timestamp
            else:
                offset = (toks[last - 2] - timestamp_begin) * input_stride  # seek to the last closed segment
        else:
            ts = [t for t in toks if t >= timestamp_begin]
            start_pos, last_pos = 0.0, seek_num_frames // 2
            skip = False
            offset = seek_num_frames
            if len(ts) > 1:
                start_pos, last_pos = ts[-2] - timestamp_begin, ts[-1] - timestamp_begin
            elif len(ts) == 1:
                start_pos = ts[-1] - timestamp_begin
                if start_pos > 200:  # the segment does not fit the window: roll back (generation.py:501-507)
                    offset = start_pos * input_stride - 100
                    skip = True
            elif len(toks) > 1:
                pass  # decoding without timestamps: keep the window as one segment
            else:
                skip = True
            if not skip:
                dur = last_pos * time_precision
                if len(ts) <= 1:  # the reference multiplies an int64 tensor here, i.e. rounds this product to fp32
                    dur = float(torch.tensor(last_pos, dtype=torch.int64) * time_precision)
                segments = [{"start": torch.tensor(time_offset + start_pos * time_precision, dtype=torch.float64),
                             "end": torch.tensor(time_offset + dur, dtype=torch.float64),
                             "tokens": seq}]
                offset = seek_num_frames
        if offset <= 0:
            raise ValueError(f"Segment offset: {offset} <= 0. This should not happen!")
        return segments, int(offset)

    # ---- re-tokenisation of global timestamps (needs the tokenizer) -----------------------------------------------
    @staticmethod
    def round_to_nearest_0_02(x) -> Decimal:
        return (Decimal(str(x)) / Decimal("0.02")).to_integral_value(rounding=ROUND_HALF_UP) * Decimal("0.02")

    def _fix_timestamps_from_segmentation(self, sequences: dict) -> torch.Tensor:
        """Fold recording-global segment times back into Whisper's 0-30 s timestamp tokens, inserting block markers
        between 30 s blocks (src/models/dicow/generation.py:322-413).  Text/tokenizer logic on the host."""
        vocab = self.tokenizer.get_vocab()
        first_ts, empty_tok = vocab["<|0.00|>"], vocab["Ġ"]
        thirty, zero = Decimal(30), Decimal(0)
        results = []
        for segs in sequences["segments"]:
            segs = [s for s in segs if len(s["tokens"]) > 0 and not (len(s["tokens"]) == 1 and int(s["tokens"][0]) == first_ts)]
            pieces = []
            prev_end = None
            corr = Decimal(0.0)
            for s in segs:
                start = self.round_to_nearest_0_02(s["start"].item())
                end = self.round_to_nearest_0_02(s["end"].item())
                toks = s["tokens"]
                block = (start + corr) // 30
                if prev_end is not None:
                    prev_block = (prev_end - Decimal("0.001")) // 30
                    if block > prev_block:
                        pieces.append((30, [empty_tok], 30))
                    for _ in range(int(block - prev_block - 1)):
                        pieces.append((0, [empty_tok], 30))
                else:
                    for _ in range(int(start // 30)):
                        pieces.append((0, [empty_tok], 30))
                if (start + corr) // 30 == (end + corr) // 30:
                    pieces.append(((start + corr) % 30, toks, (end + corr) % 30))
                elif (end + corr) % 30 == 0:
                    pieces.append(((start + corr) % 30, toks, 30))
                    corr = zero
                else:
                    new_start = (corr + start) % 30
                    new_end = (end + corr) % 30
                    if end - start == thirty:
                        if float(new_start) % 30.0 == 0.0:
                            new_end, corr = thirty, zero
                        else:
                            corr = Decimal(-0.02)
                            new_end += Decimal(corr)
                    else:
                        corr = zero
                    pieces.append((new_start, toks, new_end))
                prev_end = end + corr
            text = "".join(f"<|{a:.2f}|>{self.tokenizer.decode(t)}<|{b:.2f}|>" for a, t, b in pieces)
            results.append(self.tokenizer(text)["input_ids"])
        dev = sequences["sequences"].device
        n = max([len(r) for r in results] + [0])
        out = torch.full((len(results), n), self.tokenizer.pad_token_id, dtype=torch.int64)
        for i, r in enumerate(results):
            out[i, :len(r)] = torch.tensor(r, dtype=torch.int64)
        return out.to(dev)
