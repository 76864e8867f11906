"""Conditioning probe (CPU, fp32 oracle only -- test infrastructure): how far do the training-step gradients move when the
weight matrices are merely rounded to bf16?  A seeded draw where this alone exceeds the parity bound cannot be used to
judge a bf16 path; tests/test_gpu_training.py picks its draws with this.   python tests/cond_cpu.py [tag ...]"""
import dataclasses
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import dicow_oracle as orc  # noqa: E402
from oracle import synth  # noqa: E402

EOS, SOT, LANG, TASK, TS_BEGIN, N_TS = 257, 258, 259, 260, 262, 38
MINI = synth.Dims(**{**synth.GOLDEN_MINI.__dict__, "use_enrollments": False, "scb_layers": 0})
TINY_SHORT = dataclasses.replace(synth.WHISPER_TINY, T=200, vocab=1000, max_target=64, pad_token_id=257, eos_token_id=257,
                                 decoder_start_token_id=258, enc_layers=2, dec_layers=2)
CASES = {"mini": (MINI, 2, 11), "tiny": (TINY_SHORT, 3, 24),
         "mini-bias": (dataclasses.replace(MINI, fddt_bias_only=True), 2, 11),
         "tiny-bias": (dataclasses.replace(TINY_SHORT, fddt_bias_only=True), 3, 24),
         "mini-full": (dataclasses.replace(MINI, fddt_is_diagonal=False), 2, 11),
         "tiny-full": (dataclasses.replace(TINY_SHORT, fddt_is_diagonal=False), 3, 24),
         "mini-nots": (dataclasses.replace(MINI, remove_timestamps_from_ctc=True, vocab=1700), 2, 11),
         "tiny-l1": (dataclasses.replace(TINY_SHORT, enc_layers=3, apply_fddt_to_n_layers=1), 3, 24),
         "tiny-nopre": (dataclasses.replace(TINY_SHORT, enc_layers=3, use_pre_pos_fddt=False), 3, 24),
         "mini-layer": (dataclasses.replace(MINI, additional_layer=True), 2, 11),
         "tiny-layer": (dataclasses.replace(TINY_SHORT, additional_layer=True), 3, 24)}


def grads(dm, B, S, rounded, tag):
    p = orc.to_torch(synth.make_params(dm), device="cpu")
    if rounded:
        p = {k: (v.bfloat16().float() if v.dim() >= 2 else v) for k, v in p.items()}
    names = [k for k in p if "encoder.embed_positions" not in k]
    for n in names:
        p[n].requires_grad_(True)
    p["proj_out.weight"] = p["model.decoder.embed_tokens.weight"]
    feats = torch.from_numpy(synth.make_features(tag, B, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno(tag, B, dm.T, "soft", pad_tail=5))
    labels = torch.from_numpy(synth.make_labels(tag, B, S, min(dm.vocab, 300), EOS, TS_BEGIN, prefix=(LANG, TASK)))
    upp = labels.clone()
    upp[:, ::3] = torch.where(upp[:, ::3] >= 0, (upp[:, ::3] + 3) % 250, upp[:, ::3])
    loss, _, enc = orc.model_forward(p, dm, feats, stno, labels, upp, ctc_prefix_tokens=(SOT, LANG, TASK), ts_begin=TS_BEGIN,
                                     n_ts=N_TS)
    enc.retain_grad()
    loss.backward()
    return {n: p[n].grad for n in names if p[n].grad is not None}, enc.grad


def main():
    only = [a for a in sys.argv[1:] if a in CASES]
    tags = [a for a in sys.argv[1:] if a not in CASES] or ["tr1", "tr2"]
    for tag in tags:
        for name, (dm, B, S) in CASES.items():
            if only and name not in only:
                continue
            g0, de0 = grads(dm, B, S, False, tag)
            g1, de1 = grads(dm, B, S, True, tag)
            rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()  # noqa: E731
            rows = sorted(((rel(g1[n], g0[n]), n) for n in g0), reverse=True)
            print(f"{tag} {name}: d_enc {rel(de1, de0):.4f}; worst parameters: " + ", ".join(f"{n} {e:.4f}" for e, n in rows[:3]))


if __name__ == "__main__":
    main()
