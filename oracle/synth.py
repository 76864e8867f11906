"""Deterministic synthetic weights and inputs shared by the oracle, the golden-vector generator, the tests and bench.py.

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/README.md): nothing in the product package imports this module.

There are no checkpoints, tokenizer files or audio offline, so every parity check runs on
synthetic parameters.  Values are produced by a counter-based integer hash (splitmix64 on uint64 numpy arrays), so the
same (name, shape) gives bit-identical float32 values on any machine / numpy / torch version -- golden fixtures
therefore only need to store OUTPUTS of the reference, not the weights.

Parameter names and shapes follow the reference's state_dict; the FDDT weights/biases, LayerNorm
affine parameters and the SCB gate are drawn AWAY from their initial identities (reference FDDT at init is the identity,
SCB gate at init is 0: src/models/dicow/encoder.py:49-73, src/models/dicow/layers.py:79-93,135) so that tests
actually exercise them.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import numpy as np

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _fnv1a64(s: str) -> int:
    h = 0xCBF29CE484222325
    for ch in s.encode("utf-8"):
        h ^= ch
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(name: str, shape, stream: int = 0) -> np.ndarray:
    """float64 uniforms in [0, 1) keyed by (name, stream, flat index)."""
    n = int(np.prod(shape)) if len(shape) else 1
    seed = np.uint64(_fnv1a64(name) ^ ((stream * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF))
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) * np.uint64(0x2545F4914F6CDD1D) + seed
    z = _splitmix64(_splitmix64(ctr))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return u.reshape(shape)


def uniform(name: str, shape, lo: float, hi: float, stream: int = 0) -> np.ndarray:
    return (lo + (hi - lo) * uniform01(name, shape, stream)).astype(np.float32)


def gaussish(name: str, shape, std: float = 1.0, mean: float = 0.0) -> np.ndarray:
    """Approximately normal (Irwin-Hall, 4 uniforms): exact arithmetic only, so reproducible bit for bit."""
    s = sum(uniform01(name, shape, stream=k + 1) for k in range(4)) - 2.0
    return (mean + std * s * math.sqrt(3.0)).astype(np.float32)


# ------------------------------------------------------------------------------------------------------------------
# model dimensions
# ------------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Dims:
    """The subset of DiCoWConfig (reference src/models/dicow/config.py:6-59 + WhisperConfig) the hot path reads."""
    n_mels: int = 80
    d: int = 384
    enc_layers: int = 4
    heads: int = 6
    ffn: int = 1536
    dec_layers: int = 4
    dec_heads: int = 6
    dec_ffn: int = 1536
    vocab: int = 51865
    T: int = 1500            # max_source_positions (encoder frames); mel frames = 2 T
    max_target: int = 448
    use_fddt: bool = True
    use_pre_pos_fddt: bool = True
    non_target_fddt_value: float = 0.5
    ctc_weight: float = 0.3
    additional_self_attention_layer: bool = True
    pre_ctc_sub_sample: bool = True
    use_enrollments: bool = False
    scb_layers: int = 0
    apply_fddt_to_n_layers: int = -1
    fddt_is_diagonal: bool = True    # False: full d x d CustomLinear per class (src/models/dicow/layers.py:7-47)
    fddt_bias_only: bool = False     # True: one bias vector per class (src/models/dicow/FDDT.py:43-51)
    additional_layer: bool = False   # a whole encoder layer in front of the CTC head (encoder.py:17-18, 88-89)
    remove_timestamps_from_ctc: bool = False  # CTC targets without timestamp / task tokens (encoder.py:76, 111-113)
    pad_token_id: int = 50257
    eos_token_id: int = 50257
    decoder_start_token_id: int = 50258

    @property
    def n_fddt(self) -> int:
        return self.enc_layers if self.apply_fddt_to_n_layers == -1 else self.apply_fddt_to_n_layers

    def hf_kwargs(self) -> dict:
        """kwargs for the reference's DiCoWConfig (used only by tests/golden/make_golden.py)."""
        return dict(
            vocab_size=self.vocab, num_mel_bins=self.n_mels, d_model=self.d, encoder_layers=self.enc_layers,
            encoder_attention_heads=self.heads, decoder_layers=self.dec_layers, decoder_attention_heads=self.dec_heads,
            encoder_ffn_dim=self.ffn, decoder_ffn_dim=self.dec_ffn, max_source_positions=self.T,
            max_target_positions=self.max_target, use_fddt=self.use_fddt, use_pre_pos_fddt=self.use_pre_pos_fddt,
            fddt_is_diagonal=self.fddt_is_diagonal, fddt_bias_only=self.fddt_bias_only, additional_layer=self.additional_layer,
            remove_timestamps_from_ctc=self.remove_timestamps_from_ctc,
            non_target_fddt_value=self.non_target_fddt_value, fddt_init="suppressive",
            ctc_weight=self.ctc_weight, additional_self_attention_layer=self.additional_self_attention_layer,
            pre_ctc_sub_sample=self.pre_ctc_sub_sample, use_enrollments=self.use_enrollments,
            scb_layers=self.scb_layers if self.use_enrollments else None,
            apply_fddt_to_n_layers=self.apply_fddt_to_n_layers,
            pad_token_id=self.pad_token_id, eos_token_id=self.eos_token_id, bos_token_id=self.eos_token_id,
            decoder_start_token_id=self.decoder_start_token_id, suppress_tokens=None, begin_suppress_tokens=None,
            activation_function="gelu", dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
            encoder_layerdrop=0.0, decoder_layerdrop=0.0, scale_embedding=False,
        )


WHISPER_TINY = Dims()  # BASELINE.json configs[0]
LARGE_V3_TURBO = Dims(n_mels=128, d=1280, enc_layers=32, heads=20, ffn=5120, dec_layers=4, dec_heads=20, dec_ffn=5120,
                      vocab=51866)  # BASELINE.json configs[1..4]
# the miniature used for the committed golden vectors (tests/golden/): every feature on, odd T to exercise tails
GOLDEN_MINI = Dims(n_mels=16, d=128, enc_layers=3, heads=2, ffn=256, dec_layers=2, dec_heads=2, dec_ffn=256, vocab=300,
                   T=50, max_target=40, use_enrollments=True, scb_layers=2, pad_token_id=257, eos_token_id=257,
                   decoder_start_token_id=258)


def param_shapes(dm: Dims, decoder: bool = True) -> Dict[str, Tuple[int, ...]]:
    """state_dict names -> shapes, as enumerated from the reference model."""
    d = dm.d
    sh: Dict[str, Tuple[int, ...]] = {}

    def attn(prefix: str):
        sh[prefix + ".k_proj.weight"] = (d, d)
        sh[prefix + ".v_proj.weight"] = (d, d)
        sh[prefix + ".v_proj.bias"] = (d,)
        sh[prefix + ".q_proj.weight"] = (d, d)
        sh[prefix + ".q_proj.bias"] = (d,)
        sh[prefix + ".out_proj.weight"] = (d, d)
        sh[prefix + ".out_proj.bias"] = (d,)

    def ln(prefix: str):
        sh[prefix + ".weight"] = (d,)
        sh[prefix + ".bias"] = (d,)

    def fddt(prefix: str):
        for c in ("target", "non_target", "overlap", "silence"):
            if dm.fddt_bias_only:
                sh[f"{prefix}.{c}_linear"] = (d,)
                continue
            sh[f"{prefix}.{c}_linear.weight"] = (d,) if dm.fddt_is_diagonal else (d, d)
            sh[f"{prefix}.{c}_linear.bias"] = (d,)

    def layer(p: str):
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        sh[p + ".fc1.weight"] = (dm.ffn, d)
        sh[p + ".fc1.bias"] = (dm.ffn,)
        sh[p + ".fc2.weight"] = (d, dm.ffn)
        sh[p + ".fc2.bias"] = (d,)
        ln(p + ".final_layer_norm")

    e = "model.encoder"
    sh[e + ".conv1.weight"] = (d, dm.n_mels, 3)
    sh[e + ".conv1.bias"] = (d,)
    sh[e + ".conv2.weight"] = (d, d, 3)
    sh[e + ".conv2.bias"] = (d,)
    sh[e + ".embed_positions.weight"] = (dm.T, d)
    for i in range(dm.enc_layers):
        p = f"{e}.layers.{i}"
        attn(p + ".self_attn")
        ln(p + ".self_attn_layer_norm")
        sh[p + ".fc1.weight"] = (dm.ffn, d)
        sh[p + ".fc1.bias"] = (dm.ffn,)
        sh[p + ".fc2.weight"] = (d, dm.ffn)
        sh[p + ".fc2.bias"] = (d,)
        ln(p + ".final_layer_norm")
    ln(e + ".layer_norm")
    if dm.ctc_weight > 0:
        if dm.additional_layer:
            layer(e + ".additional_layer")
        if dm.additional_self_attention_layer:
            attn(e + ".additional_self_attention_layer")
        if dm.pre_ctc_sub_sample:
            sh[e + ".subsample_conv1.weight"] = (d, d, 3)
            sh[e + ".subsample_conv2.weight"] = (d, d, 3)
        sh[e + ".lm_head.weight"] = (dm.vocab + 1, d)
    if dm.use_fddt:
        for i in range(dm.n_fddt):
            fddt(f"{e}.fddts.{i}")
        if dm.use_pre_pos_fddt:
            fddt(e + ".initial_fddt")
    if dm.use_enrollments:
        for i in range(dm.scb_layers):
            p = f"{e}.ca_enrolls.{i}.cae"
            attn(p + ".cross_attn")
            sh[p + ".cross_gate.gate"] = (1,)
            sh[p + ".ffn.0.weight"] = (dm.ffn, 2 * d)
            sh[p + ".ffn.0.bias"] = (dm.ffn,)
            sh[p + ".ffn.3.weight"] = (d, dm.ffn)
            sh[p + ".ffn.3.bias"] = (d,)
    if decoder:
        dd = "model.decoder"
        sh[dd + ".embed_tokens.weight"] = (dm.vocab, d)
        sh[dd + ".embed_positions.weight"] = (dm.max_target, d)
        for i in range(dm.dec_layers):
            p = f"{dd}.layers.{i}"
            attn(p + ".self_attn")
            ln(p + ".self_attn_layer_norm")
            attn(p + ".encoder_attn")
            ln(p + ".encoder_attn_layer_norm")
            sh[p + ".fc1.weight"] = (dm.dec_ffn, d)
            sh[p + ".fc1.bias"] = (dm.dec_ffn,)
            sh[p + ".fc2.weight"] = (d, dm.dec_ffn)
            sh[p + ".fc2.bias"] = (d,)
            ln(p + ".final_layer_norm")
        ln(dd + ".layer_norm")
    return sh


def make_param(name: str, shape: Tuple[int, ...], seed: str = "w0") -> np.ndarray:
    key = f"{seed}/{name}"
    leaf = name.rsplit(".", 1)[-1]
    if "fddt" in name:  # perturbed off the identity (weights 1 -> U(0.5, 1.5), biases 0 -> U(-0.2, 0.2))
        if leaf == "weight" and len(shape) == 2:  # full-matrix FDDT: a perturbed identity plus a dense component
            a = math.sqrt(3.0 / shape[1]) * 0.5
            return (np.eye(shape[0], dtype=np.float32) * uniform(key + "/diag", (shape[0],), 0.5, 1.5)[:, None]
                    + uniform(key, shape, -a, a)).astype(np.float32)
        return uniform(key, shape, 0.5, 1.5) if leaf == "weight" else uniform(key, shape, -0.2, 0.2)
    if leaf == "gate":
        return np.full(shape, 0.5, np.float32)
    if "layer_norm" in name:
        return uniform(key, shape, 0.8, 1.2) if leaf == "weight" else uniform(key, shape, -0.1, 0.1)
    if "embed_positions" in name:
        return uniform(key, shape, -0.5, 0.5)
    if "embed_tokens" in name:
        return uniform(key, shape, -0.6, 0.6)
    if leaf == "bias":
        return uniform(key, shape, -0.1, 0.1)
    fan_in = int(np.prod(shape[1:]))
    a = math.sqrt(3.0 / fan_in) * (1.4 if "lm_head" in name else 1.0)
    return uniform(key, shape, -a, a)


def make_params(dm: Dims, decoder: bool = True, seed: str = "w0") -> Dict[str, np.ndarray]:
    """All parameters as float32 numpy arrays keyed by reference state_dict names (proj_out is tied to embed_tokens)."""
    out = {k: make_param(k, s, seed) for k, s in param_shapes(dm, decoder).items()}
    if decoder:
        out["proj_out.weight"] = out["model.decoder.embed_tokens.weight"]
    return out


# ------------------------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------------------------
def make_features(name: str, B: int, n_mels: int, F: int) -> np.ndarray:
    """Normalised log-mel-like features: clamp(N(-0.3, 0.4), -1, 1.5), float32 [B, n_mels, F]."""
    return np.clip(gaussish("feat/" + name, (B, n_mels, F), std=0.4, mean=-0.3), -1.0, 1.5).astype(np.float32)


def make_stno(name: str, B: int, T: int, kind: str = "soft", pad_tail: int = 0) -> np.ndarray:
    """STNO masks float32 [B, 4, T] (class order silence, target, non-target, overlap: src/models/dicow/FDDT.py:41-63).

    soft: rows of a softmax (sum to 1 per frame, like the 320-sample averages of src/data/local_datasets.py:185-194);
    hard: one-hot runs of 5..40 frames; pad_tail frames at the end are silence=1 (src/data/collators.py:157-161)."""
    if kind == "soft":
        z = 3.0 * gaussish("stno/" + name, (B, 4, T)).astype(np.float64)
        z = np.exp(z - z.max(axis=1, keepdims=True))
        m = (z / z.sum(axis=1, keepdims=True)).astype(np.float32)
    else:
        u = uniform01("stno_cls/" + name, (B, T))
        ln_ = uniform01("stno_len/" + name, (B, T))
        m = np.zeros((B, 4, T), np.float32)
        for b in range(B):
            t = 0
            k = 0
            while t < T:
                c = int(u[b, k] * 4) % 4
                n = 5 + int(ln_[b, k] * 36)
                m[b, c, t:t + n] = 1.0
                t += n
                k += 1
    if pad_tail > 0:
        m[:, :, T - pad_tail:] = 0.0
        m[:, 0, T - pad_tail:] = 1.0
    return m


def make_audio(name: str, n_samples: int) -> np.ndarray:
    """0.1 * N(0,1)-like float32 waveform with a slow amplitude envelope (so the log-mel floor is exercised)."""
    x = gaussish("wav/" + name, (n_samples,), std=0.1)
    t = np.arange(n_samples, dtype=np.float64) / 16000.0
    env = (0.55 + 0.45 * np.sin(2.0 * np.pi * 0.37 * t)).astype(np.float32)
    return (x * env).astype(np.float32)


def make_labels(name: str, B: int, S: int, vocab: int, eos: int, ts_begin: Optional[int] = None,
                prefix: Tuple[int, ...] = ()) -> np.ndarray:
    """int64 [B, S] label rows: prefix, <|0.00|>-style timestamp, text ids, timestamp, eos, then -100 padding."""
    lab = np.full((B, S), -100, np.int64)
    u = uniform01("lab/" + name, (B, S))
    ln_ = uniform01("lablen/" + name, (B,))
    n_text_ids = min(vocab, eos) if ts_begin is None else min(eos, ts_begin)
    for b in range(B):
        n = max(len(prefix) + 4, int(S * (0.5 + 0.5 * ln_[b])))
        row = list(prefix)
        if ts_begin is not None:
            row.append(ts_begin)
        while len(row) < n - (2 if ts_begin is not None else 1):
            row.append(int(u[b, len(row)] * n_text_ids) % n_text_ids)
        if ts_begin is not None:
            row.append(min(vocab - 1, ts_begin + 7 + b))
        row.append(eos)
        lab[b, :len(row)] = row[:S]
    return lab
