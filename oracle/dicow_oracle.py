"""CPU oracle: a plain-PyTorch fp32 restatement of the reference's hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file; the
product package (ts-asr-whisper_b200/) never does and fails loudly without its CUDA library.

The reference (BUTSpeechFIT/TS-ASR-Whisper, /root/reference) is Python over HF transformers; its arithmetic lives partly
in its own src/models/dicow/*.py and partly in the third-party `transformers` package (pinned 4.55.0 in
requirements.txt:22, installed here: 5.5.0; module transformers.models.whisper.{modeling_whisper,
feature_extraction_whisper}, transformers.generation.logits_process).  Each function below cites the file:line it
restates ("HF:" = transformers/).  Parity pinning: the reference ships no tests and no golden vectors (SURVEY.md
section 4), so this oracle is pinned against outputs of the reference itself, executed in the build container by
tests/golden/make_golden.py (reference modules imported from /root/reference/src + the installed transformers) and
committed as tests/golden/*.npz; tests/test_oracle_golden.py replays them on CPU.

Everything is a pure function of (params: dict name -> fp32 tensor keyed by the reference's state_dict names, dims,
inputs).  fp32, eval mode (all dropouts are identity, LayerDrop off).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .synth import Dims

Params = Dict[str, torch.Tensor]


def to_torch(params_np: Dict[str, np.ndarray], device="cpu", dtype=torch.float32) -> Params:
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device=device, dtype=dtype) for k, v in params_np.items()}


# ----------------------------------------------------------------------------------------------------------------
# A1  log-mel front-end
# ----------------------------------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f: np.ndarray) -> np.ndarray:
    # HF:audio_utils.py hertz_to_mel(mel_scale="slaney"): linear below 1 kHz, log above
    f = np.asarray(f, dtype=np.float64)
    mel = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    hi = f >= 1000.0
    mel = np.where(hi, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * logstep, mel)
    return mel


def _mel_to_hz_slaney(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    hi = m >= 15.0
    return np.where(hi, 1000.0 * np.exp(logstep * (m - 15.0)), f)


def mel_filterbank(n_mels: int, n_fft: int = 400, sr: int = 16000, fmax: float = 8000.0) -> np.ndarray:
    """[n_fft//2+1, n_mels] float32 slaney-normalised triangular filters.

    Restates HF:audio_utils.py mel_filter_bank(num_frequency_bins=201, min 0, max 8000, norm="slaney",
    mel_scale="slaney") as constructed at HF:models/whisper/feature_extraction_whisper.py:95-103."""
    nfreq = 1 + n_fft // 2
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(fmax), n_mels + 2)
    filter_freqs = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, sr // 2, nfreq)
    fdiff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / fdiff[:-1]
    up = slopes[:, 2:] / fdiff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2:n_mels + 2] - filter_freqs[:n_mels])
    fb = fb * np.expand_dims(enorm, 0)
    return fb.astype(np.float32)


def log_mel(wave: np.ndarray, n_mels: int, chunk_samples: int = 480000, n_fft: int = 400, hop: int = 160
            ) -> Tuple[np.ndarray, np.ndarray]:
    """One recording -> (input_features [n_mels, frames] fp32, attention_mask [frames] int32).

    Restates the reference call site src/data/local_datasets.py:208-214 (padding="longest",
    pad_to_multiple_of=n_samples, truncation=False, return_attention_mask=True) over
    HF:models/whisper/feature_extraction_whisper.py:135-164 (_torch_extract_fbank_features) and :328-337 (mask
    rescale): zero-pad to a multiple of 30 s, reflect-padded centred STFT with a periodic Hann window, drop the last
    frame, power, mel, log10 clamp 1e-10, floor at (global max - 8), (x + 4) / 4."""
    wave = np.asarray(wave, dtype=np.float32)
    n = wave.shape[0]
    n_pad = ((n + chunk_samples - 1) // chunk_samples) * chunk_samples
    x = np.zeros(n_pad, np.float32)
    x[:n] = wave
    mask = np.zeros(n_pad, np.int32)
    mask[:n] = 1
    xt = torch.from_numpy(x)
    window = torch.hann_window(n_fft)
    stft = torch.stft(xt, n_fft, hop, window=window, return_complex=True)
    mag = stft[..., :-1].abs() ** 2
    fb = torch.from_numpy(mel_filterbank(n_mels, n_fft))
    mel = fb.T @ mag
    log_spec = torch.clamp(mel, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec.numpy(), mask[::hop].copy()


# ----------------------------------------------------------------------------------------------------------------
# A2  STNO mask from per-speaker sample-level activity (src/data/local_datasets.py:162-196)
# ----------------------------------------------------------------------------------------------------------------
def stno_mask(activity: np.ndarray, target: int, window_samples: int = 480000, frame_samples: int = 320) -> np.ndarray:
    """activity [n_speakers, n_samples] bool, ``target`` = row of the target speaker (-1: none of them).  Zero-pad to whole
    30 s windows, a_i[t] = mean activity of speaker i over the 320 samples of encoder frame t, then per frame
      S = prod_i (1 - a_i),  T = a_s prod_{i != s} (1 - a_i),  N = (1 - a_s)(1 - prod_{i != s} (1 - a_i)),  O = a_s - T
    in float32, speakers multiplied in row order.  Returns [frames, 4] (S, T, N, O) like the reference."""
    n_spk, n = activity.shape
    pad = (window_samples - n) % window_samples
    frames = (n + pad) // frame_samples
    a = np.zeros((n_spk, frames), np.float32)
    for i in range(n_spk):
        row = np.zeros(n + pad, np.float32)
        row[:n] = activity[i]
        a[i] = row.reshape(frames, frame_samples).sum(axis=1) / np.float32(frame_samples)
    out = np.zeros((frames, 4), np.float32)
    one = np.float32(1.0)
    sil = np.ones(frames, np.float32)
    others = np.ones(frames, np.float32)
    for i in range(n_spk):
        sil = sil * (one - a[i])
        if i != target:
            others = others * (one - a[i])
    a_s = a[target] if target >= 0 else np.zeros(frames, np.float32)
    tgt = a_s * others
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = sil, tgt, (one - a_s) * (one - others), a_s - tgt
    return out


# ----------------------------------------------------------------------------------------------------------------
# A4/A5  FDDT
# ----------------------------------------------------------------------------------------------------------------
_FDDT_ORDER = ("silence", "target", "non_target", "overlap")  # STNO mask channel order, FDDT.py:56-62


def fddt_tables(p: Params, prefix: str) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """weights ([4, d] diagonal, [4, d, d] full-matrix, None for the bias-only variant) and biases [4, d] in STNO order"""
    if f"{prefix}.{_FDDT_ORDER[0]}_linear" in p:  # bias-only: the parameter IS the bias vector (FDDT.py:10,43-51)
        return None, torch.stack([p[f"{prefix}.{c}_linear"] for c in _FDDT_ORDER])
    w = torch.stack([p[f"{prefix}.{c}_linear.weight"] for c in _FDDT_ORDER])
    b = torch.stack([p[f"{prefix}.{c}_linear.bias"] for c in _FDDT_ORDER])
    return w, b


def fddt(x: torch.Tensor, stno: torch.Tensor, w: Optional[torch.Tensor], b: torch.Tensor) -> torch.Tensor:
    """src/models/dicow/FDDT.py:41-63.  Default: sum_c stno[:, c, :, None] * lin_c(x) with lin_c diagonal (w_c * x + b_c,
    layers.py:73-77) or a full nn.Linear (x W_c^T + b_c, layers.py:7-47); bias-only: x + sum_c stno_c * b_c.
    x [B, T, d], stno [B, 4, T]."""
    if w is None:
        out = x
        for c in range(4):
            out = out + stno[:, c, :, None] * b[c]
        return out
    out = torch.zeros_like(x)
    for c in range(4):
        y = F.linear(x, w[c], b[c]) if w.dim() == 3 else x * w[c] + b[c]
        out = out + y * stno[:, c, :, None]
    return out


# ----------------------------------------------------------------------------------------------------------------
# A8  attention / encoder layer (third-party HF:models/whisper/modeling_whisper.py)
# ----------------------------------------------------------------------------------------------------------------
def lora_linear(p: Params, name: str, x: torch.Tensor, bias: bool = True) -> torch.Tensor:
    """nn.Linear ``name``, with a LoRA adapter when the dict holds one: y = x W^T + b + (alpha / r) * (x A^T) B^T
    (peft.tuners.lora.layer.Linear.forward, third-party peft -- not in this image; the reference's use:
    src/models/containers.py:69-78, r = 16, lora_alpha = 32, dropout 0, bias "none", decoder q/k/v/out_proj/fc1/fc2).
    Keys: ``name + ".lora_A"`` [r, in], ``name + ".lora_B"`` [out, r], optional ``"lora_scale"`` (default alpha / r = 2)."""
    y = F.linear(x, p[name + ".weight"], p.get(name + ".bias") if bias else None)
    if (name + ".lora_A") in p:
        y = y + float(p.get("lora_scale", 2.0)) * F.linear(F.linear(x, p[name + ".lora_A"]), p[name + ".lora_B"])
    return y


def attention(p: Params, prefix: str, x_q: torch.Tensor, x_kv: torch.Tensor, heads: int, causal: bool = False
              ) -> torch.Tensor:
    """HF:modeling_whisper.py:284-357: q = (x Wq + bq) * hd^-0.5 scaled BEFORE QK^T, k has no bias, softmax over
    keys with scaling 1.0, no padding mask, out_proj with bias."""
    B, Tq, d = x_q.shape
    Tk = x_kv.shape[1]
    hd = d // heads
    q = lora_linear(p, prefix + ".q_proj", x_q) * (hd ** -0.5)
    k = lora_linear(p, prefix + ".k_proj", x_kv, bias=False)
    v = lora_linear(p, prefix + ".v_proj", x_kv)
    q = q.view(B, Tq, heads, hd).transpose(1, 2)
    k = k.view(B, Tk, heads, hd).transpose(1, 2)
    v = v.view(B, Tk, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, is_causal=causal, scale=1.0)
    o = o.transpose(1, 2).reshape(B, Tq, d)
    return lora_linear(p, prefix + ".out_proj", o)


def layer_norm(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), p[prefix + ".weight"], p[prefix + ".bias"], 1e-5)


def encoder_layer(p: Params, prefix: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """HF:modeling_whisper.py:380-414 (pre-LN block, exact-erf GELU, dropout = identity in eval)."""
    x = x + attention(p, prefix + ".self_attn", (h := layer_norm(p, prefix + ".self_attn_layer_norm", x)), h, heads)
    h = layer_norm(p, prefix + ".final_layer_norm", x)
    h = F.gelu(F.linear(h, p[prefix + ".fc1.weight"], p[prefix + ".fc1.bias"]))
    return x + F.linear(h, p[prefix + ".fc2.weight"], p[prefix + ".fc2.bias"])


# ----------------------------------------------------------------------------------------------------------------
# A6  SE-DiCoW speaker communication block
# ----------------------------------------------------------------------------------------------------------------
def scb(p: Params, prefix: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """src/models/dicow/layers.py:145-170,180-193.  x [2B, T, d] with rows interleaved target/enrollment.
    q = target stream, kv = enrollment stream (no LayerNorm); ffn(cat[attn, q]); q + tanh(gate) * upd; the
    enrollment stream is returned unchanged."""
    B2, T, d = x.shape
    xr = x.view(B2 // 2, 2, T, d)
    q, kv = xr[:, 0], xr[:, 1]
    a = attention(p, prefix + ".cae.cross_attn", q, kv, heads)
    cat = torch.cat([a, q], dim=-1)
    h = F.gelu(F.linear(cat, p[prefix + ".cae.ffn.0.weight"], p[prefix + ".cae.ffn.0.bias"]))
    upd = F.linear(h, p[prefix + ".cae.ffn.3.weight"], p[prefix + ".cae.ffn.3.bias"])
    q_out = q + torch.tanh(p[prefix + ".cae.cross_gate.gate"]) * upd
    return torch.stack([q_out, kv], dim=1).view(B2, T, d)


# ----------------------------------------------------------------------------------------------------------------
# A7  DiCoWEncoder.forward
# ----------------------------------------------------------------------------------------------------------------
def encoder_stem(p: Params, dm: Dims, input_features: torch.Tensor, stno: Optional[torch.Tensor]) -> torch.Tensor:
    """conv1+GELU, conv2(stride 2)+GELU, [B,T,d], initial FDDT, + positions: src/models/dicow/encoder.py:167-179."""
    e = "model.encoder"
    x = F.gelu(F.conv1d(input_features, p[e + ".conv1.weight"], p[e + ".conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, p[e + ".conv2.weight"], p[e + ".conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1)
    if dm.use_fddt and dm.use_pre_pos_fddt:
        x = fddt(x, stno, *fddt_tables(p, e + ".initial_fddt"))
    return x + p[e + ".embed_positions.weight"]


def encoder_forward(p: Params, dm: Dims, input_features: torch.Tensor, stno_mask: Optional[torch.Tensor] = None,
                    enrollments: Optional[dict] = None, return_logits: bool = False,
                    collect: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
    """src/models/dicow/encoder.py:140-246.  Returns last_hidden_state [B, T, d] or CTC logits [B, T/4, V+1]."""
    e = "model.encoder"
    if enrollments is not None:  # encoder.py:152-154: interleave target / enrollment on the batch axis
        input_features = torch.stack((input_features, enrollments["input_features"]), dim=1).flatten(0, 1)
        stno_mask = torch.stack((stno_mask, enrollments["stno_mask"]), dim=1).flatten(0, 1)
    if input_features.shape[-1] != 2 * dm.T:  # encoder.py:156-160
        raise ValueError(f"Whisper expects the mel input features to be of length {2 * dm.T}, "
                         f"but found {input_features.shape[-1]}.")
    x = encoder_stem(p, dm, input_features, stno_mask)
    use_enr = dm.use_enrollments and enrollments is not None
    for i in range(dm.enc_layers):
        if dm.use_fddt and i < dm.n_fddt:  # encoder.py:205-206
            x = fddt(x, stno_mask, *fddt_tables(p, f"{e}.fddts.{i}"))
        if dm.use_enrollments and i < dm.scb_layers:  # encoder.py:208-213
            x = scb(p, f"{e}.ca_enrolls.{i}", x, dm.heads)
            if i == dm.scb_layers - 1:
                x = x[::2]
                stno_mask = stno_mask[::2]
        x = encoder_layer(p, f"{e}.layers.{i}", x, dm.heads)
        if collect is not None:
            collect.append(x)
    del use_enr
    x = layer_norm(p, e + ".layer_norm", x)
    if return_logits:
        return ctc_logits(p, dm, x)
    return x


# ----------------------------------------------------------------------------------------------------------------
# A9  CTC head and loss
# ----------------------------------------------------------------------------------------------------------------
def ctc_logits(p: Params, dm: Dims, h: torch.Tensor) -> torch.Tensor:
    """possibly_update_last_hidden_states + lm_head: src/models/dicow/encoder.py:87-106,236.  The attention output
    REPLACES the hidden state (no residual, no LN); two stride-2 convs without bias; lm_head without bias."""
    e = "model.encoder"
    if dm.additional_layer:  # encoder.py:88-89: a whole WhisperEncoderLayer
        h = encoder_layer(p, e + ".additional_layer", h, dm.heads)
    elif dm.additional_self_attention_layer:
        h = attention(p, e + ".additional_self_attention_layer", h, h, dm.heads)
    if dm.pre_ctc_sub_sample:
        h = h.transpose(1, 2)
        h = F.conv1d(h, p[e + ".subsample_conv1.weight"], None, stride=2, padding=1)
        h = F.conv1d(h, p[e + ".subsample_conv2.weight"], None, stride=2, padding=1)
        h = h.transpose(1, 2)
    return F.linear(h, p[e + ".lm_head.weight"])


def ctc_label_filter(labels: torch.Tensor, dm: Dims) -> torch.Tensor:
    """encoder.py:76,111-113: with ``remove_timestamps_from_ctc`` the CTC targets keep only ids below
    vocab - 30 * 50 - 1 - 6 (the first task token; padding -100 stays), re-padded with -100 to the longest row"""
    if not dm.remove_timestamps_from_ctc:
        return labels
    first_task_token = dm.vocab - 30 * 50 - 1 - 6
    rows = [[int(v) for v in row if int(v) < first_task_token] for row in labels]
    n = max(len(r) for r in rows)
    return torch.tensor([r + [-100] * (n - len(r)) for r in rows], dtype=labels.dtype, device=labels.device).reshape(len(rows), n)


def ctc_loss(logits: torch.Tensor, labels: torch.Tensor, reduction: str = "mean") -> torch.Tensor:
    """src/models/dicow/encoder.py:108-135: fp32 log-softmax, blank = last class, input length = all frames,
    targets = labels >= 0, zero_infinity=True."""
    B, Tp, _ = logits.shape
    input_lengths = torch.full((B,), Tp, dtype=torch.long)
    target_lengths = (labels >= 0).sum(-1)
    lp = F.log_softmax(logits.float(), dim=-1).transpose(0, 1)
    return F.ctc_loss(lp, labels, input_lengths, target_lengths, blank=logits.shape[-1] - 1, reduction=reduction,
                      zero_infinity=True)


# ----------------------------------------------------------------------------------------------------------------
# A10  decoder (third-party HF:models/whisper/modeling_whisper.py:449-506, 691-796)
# ----------------------------------------------------------------------------------------------------------------
def decoder_forward(p: Params, dm: Dims, input_ids: torch.Tensor, enc: torch.Tensor, past_len: int = 0
                    ) -> torch.Tensor:
    """Teacher-forced / full-prefix decoder: embed + learned positions, dec_layers x [causal self-attn, cross-attn,
    MLP] (pre-LN), final LN.  Returns hidden [B, S, d].  (Recomputes the whole prefix: the KV cache of the reference
    is an optimisation with identical results.)"""
    dd = "model.decoder"
    S = input_ids.shape[1]
    x = p[dd + ".embed_tokens.weight"][input_ids] + p[dd + ".embed_positions.weight"][past_len:past_len + S]
    for i in range(dm.dec_layers):
        pre = f"{dd}.layers.{i}"
        h = layer_norm(p, pre + ".self_attn_layer_norm", x)
        x = x + attention(p, pre + ".self_attn", h, h, dm.dec_heads, causal=S > 1)
        h = layer_norm(p, pre + ".encoder_attn_layer_norm", x)
        x = x + attention(p, pre + ".encoder_attn", h, enc, dm.dec_heads)
        h = layer_norm(p, pre + ".final_layer_norm", x)
        h = F.gelu(lora_linear(p, pre + ".fc1", h))
        x = x + lora_linear(p, pre + ".fc2", h)
    return layer_norm(p, dd + ".layer_norm", x)


def shift_tokens_right(labels: torch.Tensor, pad_id: int, start_id: int) -> torch.Tensor:
    """HF:models/whisper/modeling_whisper.py shift_tokens_right (call site src/models/dicow/modeling_dicow.py:275-279)."""
    out = labels.new_zeros(labels.shape)
    out[:, 1:] = labels[:, :-1].clone()
    out[:, 0] = start_id
    out.masked_fill_(out == -100, pad_id)
    return out


def timestamp_smoothing(n_ts: int, sigma: float = 0.08, step: float = 0.02) -> torch.Tensor:
    """[n_ts, n_ts] row-normalised Gaussian over timestamp ids: src/models/dicow/modeling_dicow.py:56-70."""
    t = torch.arange(n_ts, dtype=torch.float32) * step
    w = torch.exp(-((t[:, None] - t[None, :]) ** 2) / (2 * sigma ** 2))
    return w / w.sum(dim=1, keepdim=True)


def decoder_loss(logits: torch.Tensor, labels: torch.Tensor, upp_labels: Optional[torch.Tensor],
                 ts_begin: Optional[int] = None, n_ts: int = 1501) -> torch.Tensor:
    """Soft-label CE of src/models/dicow/modeling_dicow.py:95-144 (ts_begin given: timestamp rows are Gaussian
    smoothed, lower/upper-case streams, per-token min, mean over non-pad) or, with ts_begin None, the hard-label
    fallback of modeling_dicow.py:312-323 (no tokenizer: mean over ALL positions, -100 ignored by CE -> 0)."""
    V = logits.shape[-1]
    flat = logits.reshape(-1, V).float()
    if ts_begin is None:
        l1 = F.cross_entropy(flat, labels.reshape(-1), reduction="none")
        if upp_labels is None:
            return l1.mean()
        l2 = F.cross_entropy(flat, upp_labels.reshape(-1), reduction="none")
        return torch.minimum(l1, l2).mean()
    lsm = F.log_softmax(flat, dim=-1)
    smooth = timestamp_smoothing(n_ts).to(flat.device)

    def soft_ce(lab: torch.Tensor) -> torch.Tensor:
        lab = lab.reshape(-1)
        tgt = F.one_hot(lab.clamp(min=0), V).float()
        is_ts = (lab >= ts_begin) & (lab < ts_begin + n_ts)
        if is_ts.any():
            rows = torch.zeros(int(is_ts.sum()), V, device=flat.device)
            rows[:, ts_begin:ts_begin + n_ts] = smooth[lab[is_ts] - ts_begin]
            tgt[is_ts] = rows
        return -(tgt * lsm).sum(-1)

    mask = (labels.reshape(-1) != -100).float()
    lo = soft_ce(labels) * mask
    up = soft_ce(upp_labels) * mask if upp_labels is not None else lo
    return torch.minimum(lo, up).sum() / mask.sum().clamp(min=1)


def model_forward(p: Params, dm: Dims, input_features, stno_mask, labels, upp_labels=None, enrollments=None,
                  ctc_prefix_tokens: Sequence[int] = (), ts_begin: Optional[int] = None, n_ts: int = 1501
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """DiCoWForConditionalGeneration.forward with labels: src/models/dicow/modeling_dicow.py:248-354.
    Returns (loss, decoder logits, encoder_last_hidden_state)."""
    enc = encoder_forward(p, dm, input_features, stno_mask, enrollments)
    dec_in = shift_tokens_right(labels, dm.pad_token_id, dm.decoder_start_token_id)
    hid = decoder_forward(p, dm, dec_in, enc)
    logits = F.linear(hid, p["proj_out.weight"])
    dec_loss = decoder_loss(logits, labels, upp_labels, ts_begin, n_ts)
    if dm.ctc_weight > 0:
        enc_logits = ctc_logits(p, dm, enc)
        enc_labels = labels.clone()
        for tok in ctc_prefix_tokens:  # modeling_dicow.py:330-332
            if bool((enc_labels[:, 0] == tok).all()):
                enc_labels = enc_labels[:, 1:]
        enc_labels[enc_labels == dm.eos_token_id] = -100
        loss = (1 - dm.ctc_weight) * dec_loss + dm.ctc_weight * ctc_loss(enc_logits, ctc_label_filter(enc_labels, dm))
    else:
        loss = dec_loss
    return loss, logits, enc


# ----------------------------------------------------------------------------------------------------------------
# A13/A14  greedy decoding with the Whisper timestamp rules
# ----------------------------------------------------------------------------------------------------------------
def timestamp_rules(input_ids: torch.Tensor, scores: torch.Tensor, *, begin_index: int, eos: int, no_timestamps: int,
                    ts_begin: int, max_initial_timestamp_index: Optional[int] = None) -> torch.Tensor:
    """WhisperTimeStampLogitsProcessor (third-party HF:generation/logits_process.py:1905-2043) followed by the DiCoW
    override restoring the EOS logit at the first generated position (src/models/dicow/utils.py:5-14).
    scores fp32 [B, V] (already suppress-token processed); returns processed copy."""
    orig = scores
    s = scores.clone()
    s[:, no_timestamps] = -float("inf")
    for k in range(input_ids.shape[0]):
        seq = input_ids[k, begin_index:].tolist()
        last_was_ts = len(seq) >= 1 and seq[-1] >= ts_begin
        penult_was_ts = len(seq) < 2 or seq[-2] >= ts_begin
        if last_was_ts:
            if penult_was_ts:
                s[k, ts_begin:] = -float("inf")  # has to be non-timestamp
            else:
                s[k, :eos] = -float("inf")  # cannot be normal text tokens
        ts = [t for t in seq if t >= ts_begin]
        if len(ts) > 0:
            # timestamps shouldn't decrease; forbid timestamp tokens smaller than the last
            if last_was_ts and not penult_was_ts:
                ts_last = ts[-1]
            else:
                ts_last = ts[-1] + 1  # avoid back-to-back identical single timestamps
            s[k, ts_begin:ts_last] = -float("inf")
    if input_ids.shape[1] == begin_index:
        s[:, :ts_begin] = -float("inf")
        if max_initial_timestamp_index is not None:
            s[:, ts_begin + max_initial_timestamp_index + 1:] = -float("inf")
    logprobs = F.log_softmax(s.float(), dim=-1)
    for k in range(input_ids.shape[0]):
        ts_lp = torch.logsumexp(logprobs[k, ts_begin:], dim=-1)
        max_text = logprobs[k, :ts_begin].max()
        if ts_lp > max_text:
            s[k, :ts_begin] = -float("inf")
    if input_ids.shape[1] == begin_index:  # src/models/dicow/utils.py:10-12
        s[:, eos] = orig[:, eos]
    return s


def greedy_decode(p: Params, dm: Dims, enc: torch.Tensor, prompt: torch.Tensor, max_new_tokens: int, *,
                  suppress: Sequence[int], no_timestamps: int, ts_begin: int,
                  return_logits: bool = False, timestamps: bool = True):
    """Greedy branch of DiCoWGenerationMixin._sample (src/models/dicow/generation.py:707-782): per step
    logits[:, -1].float() -> SuppressTokensLogitsProcessor -> timestamp processor -> argmax; finished rows emit pad
    (= eos); stop when all rows are finished or max_new_tokens reached.  Returns token ids [B, prompt+n]."""
    B = enc.shape[0]
    ids = prompt.clone()
    begin_index = prompt.shape[1]
    unfinished = torch.ones(B, dtype=torch.bool)
    sup = torch.tensor(list(suppress), dtype=torch.long)
    all_logits = []
    for _ in range(max_new_tokens):
        hid = decoder_forward(p, dm, ids, enc)
        logits = F.linear(hid[:, -1], p["proj_out.weight"]).float()
        if return_logits:
            all_logits.append(logits.clone())
        if sup.numel():
            logits[:, sup] = -float("inf")
        if timestamps:  # return_timestamps=False: HF adds no timestamp processor (generation_whisper.py _retrieve_logit_processors)
            logits = timestamp_rules(ids, logits, begin_index=begin_index, eos=dm.eos_token_id,
                                     no_timestamps=no_timestamps, ts_begin=ts_begin)
        nxt = torch.argmax(logits, dim=-1)
        nxt = torch.where(unfinished, nxt, torch.full_like(nxt, dm.pad_token_id))
        ids = torch.cat([ids, nxt[:, None]], dim=1)
        unfinished = unfinished & (nxt != dm.eos_token_id)
        if not bool(unfinished.any()):
            break
    return (ids, all_logits) if return_logits else ids
