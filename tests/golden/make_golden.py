"""Generate the committed golden vectors by running the REFERENCE ITSELF in the build container.

    python tests/golden/make_golden.py          # needs /root/reference (read-only) -- NOT available on the GPU box

The reference ships no tests / golden vectors, so parity is pinned on outputs of its own modules:
  * src/models/dicow/{config,encoder,modeling_dicow,FDDT,layers,utils}.py imported from /root/reference/src,
  * on the installed transformers 5.5.0 (reference pins 4.55.0) with the out-of-tree compatibility shim of
    SURVEY.md section 8c (WhisperEncoderLayer.forward returns a bare tensor in 5.x; the reference indexes [0]),
  * the installed WhisperFeatureExtractor called exactly as src/data/local_datasets.py:208-214.
Weights and inputs come from oracle/synth.py (hash-based, bit-reproducible), so only OUTPUTS are stored.
Nothing here is imported by the product or by tests at run time; tests read the .npz files only.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import transformers.models.whisper.modeling_whisper as mw  # noqa: E402

_orig_layer_fwd = mw.WhisperEncoderLayer.forward


def _layer_fwd_tuple(self, hidden_states, attention_mask=None, layer_head_mask=None, output_attentions=False, **kw):
    return (_orig_layer_fwd(self, hidden_states, attention_mask, **kw),)


mw.WhisperEncoderLayer.forward = _layer_fwd_tuple  # shim #1 (4.55 tuple return)

from models.dicow.config import DiCoWConfig  # noqa: E402
from models.dicow.modeling_dicow import DiCoWForConditionalGeneration  # noqa: E402
from models.dicow.utils import WhisperTimeStampLogitsProcessorCustom  # noqa: E402
from transformers import WhisperFeatureExtractor  # noqa: E402
from transformers.modeling_outputs import BaseModelOutput  # noqa: E402
from transformers.generation.logits_process import SuppressTokensLogitsProcessor  # noqa: E402

from oracle import synth  # noqa: E402


def build_reference(dm: synth.Dims):
    cfg = DiCoWConfig(**dm.hf_kwargs())
    cfg._attn_implementation = "sdpa"
    model = DiCoWForConditionalGeneration(cfg).eval()
    params = synth.make_params(dm)
    sd = {k: torch.from_numpy(v) for k, v in params.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k == "proj_out.weight" for k in missing), missing
    model.proj_out.weight = model.model.decoder.embed_tokens.weight  # tied (train.py:109-113)
    return model


class FakeTokenizer:
    """get_vocab()/prefix_tokens only -- what SoftLabelCreator (modeling_dicow.py:37-70) and forward (:330) read."""

    def __init__(self, vocab_size: int, ts_begin: int, n_ts: int, prefix_tokens):
        self._v = {f"tok{i}": i for i in range(vocab_size)}
        for k in range(n_ts):
            del self._v[f"tok{ts_begin + k}"]
            self._v[f"<|{k * 0.02:.2f}|>"] = ts_begin + k
        self.prefix_tokens = list(prefix_tokens)

    def get_vocab(self):
        return dict(self._v)


# layout of the miniature vocabulary (mirrors export_sources/generation_config.json at small scale)
MINI = synth.GOLDEN_MINI
EOS, SOT, LANG, TASK, NOTS, TS_BEGIN, N_TS = 257, 258, 259, 260, 261, 262, 38  # 262..299 timestamps
SUPPRESS = [1, 2, 7, 8, 9, 10, 14, 25, 258, 259, 260]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    dm = MINI
    model = build_reference(dm)
    out = {}
    B = 2
    feats = torch.from_numpy(synth.make_features("g0", B, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno("g0", B, dm.T, "soft", pad_tail=7))
    enr = {"input_features": torch.from_numpy(synth.make_features("g0e", B, dm.n_mels, 2 * dm.T)),
           "stno_mask": torch.from_numpy(synth.make_stno("g0e", B, dm.T, "hard"))}
    enc = model.get_encoder()
    with torch.no_grad():
        # (1) SE-DiCoW encoder (FDDT + SCB + drop of the enrollment stream)
        out["enc_se"] = enc(feats, stno_mask=stno, enrollments=enr).last_hidden_state.numpy()
        out["ctc_logits_se"] = enc(feats, stno_mask=stno, enrollments=enr, return_logits=True).logits.numpy()
        # (2) DiCoW encoder without enrollments: build a config without SCB, same weights otherwise
        dm2 = synth.Dims(**{**dm.__dict__, "use_enrollments": False, "scb_layers": 0})
        model2 = build_reference(dm2)
        out["enc_plain"] = model2.get_encoder()(feats, stno_mask=stno).last_hidden_state.numpy()
        # (3) full forward + loss, hard-label fallback (no tokenizer): modeling_dicow.py:312-323
        labels = torch.from_numpy(synth.make_labels("g0", B, 12, dm.vocab, EOS, TS_BEGIN, prefix=(LANG, TASK)))
        upp = labels.clone()
        upp[:, 4] = (upp[:, 4] + 3) % 250
        model2.tokenizer = types.SimpleNamespace(prefix_tokens=[SOT, LANG, TASK])
        o = model2(input_features=feats, stno_mask=stno, labels=labels, upp_labels=upp)
        out["fwd_hard_loss"] = o.loss.numpy()
        out["fwd_logits"] = o.logits.numpy()
        # (4) soft-label loss (timestamp smoothing + min over case streams): modeling_dicow.py:95-144
        tok = FakeTokenizer(dm.vocab, TS_BEGIN, N_TS, [SOT, LANG, TASK])
        model2.set_tokenizer(tok)
        o = model2(input_features=feats, stno_mask=stno, labels=labels, upp_labels=upp)
        out["fwd_soft_loss"] = o.loss.numpy()
        out["labels"] = labels.numpy()
        out["upp_labels"] = upp.numpy()
        # (5) greedy decode: reference forward + the reference's logits processors, loop restated from
        #     src/models/dicow/generation.py:707-782 (the stock generate() needs 4.55 private APIs)
        gcfg = types.SimpleNamespace(no_timestamps_token_id=NOTS, eos_token_id=EOS, bos_token_id=EOS,
                                     max_initial_timestamp_index=None, _detect_timestamp_from_logprob=True)
        procs = [SuppressTokensLogitsProcessor(SUPPRESS), WhisperTimeStampLogitsProcessorCustom(gcfg, begin_index=3)]
        enc_h = model2.get_encoder()(feats, stno_mask=stno).last_hidden_state
        ids = torch.tensor([[SOT, LANG, TASK]] * B)
        unfinished = torch.ones(B, dtype=torch.bool)
        step_logits = []
        for _ in range(24):
            o = model2(encoder_outputs=BaseModelOutput(last_hidden_state=enc_h), decoder_input_ids=ids, use_cache=False)
            sc = o.logits[:, -1].float()
            step_logits.append(sc.numpy().copy())
            for pr in procs:
                sc = pr(ids, sc)
            nxt = sc.argmax(-1)
            nxt = torch.where(unfinished, nxt, torch.full_like(nxt, EOS))
            ids = torch.cat([ids, nxt[:, None]], 1)
            unfinished &= nxt != EOS
            if not unfinished.any():
                break
        out["greedy_ids"] = ids.numpy()
        out["greedy_first_logits"] = step_logits[0]
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})
    print("greedy ids:", out["greedy_ids"].tolist())
    np.savez_compressed(os.path.join(HERE, "mini_model.npz"), **out)

    # (6) log-mel: installed WhisperFeatureExtractor exactly as src/data/local_datasets.py:208-214,
    #     at chunk_length=2 s (n_samples=32000) to keep the fixture small; one 1.3-window recording => 2 windows share a floor
    mel_out = {}
    for n_mels in (80, 128):
        fe = WhisperFeatureExtractor(feature_size=n_mels, chunk_length=2)
        wav = synth.make_audio(f"mel{n_mels}", 41777)
        f = fe(wav, return_tensors="pt", sampling_rate=16000, return_attention_mask=True, truncation=False,
               padding="longest", pad_to_multiple_of=fe.n_samples)
        mel_out[f"feat{n_mels}"] = f.input_features[0].numpy()
        mel_out[f"mask{n_mels}"] = f.attention_mask[0].numpy().astype(np.int32)
    fe = WhisperFeatureExtractor(feature_size=128)
    mel_out["filters128"] = fe.mel_filters.astype(np.float32)
    mel_out["filters80"] = WhisperFeatureExtractor(feature_size=80).mel_filters.astype(np.float32)
    print({k: v.shape for k, v in mel_out.items()})
    np.savez_compressed(os.path.join(HERE, "mel.npz"), **mel_out)


if __name__ == "__main__":
    main()
