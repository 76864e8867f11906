"""Golden vectors for joint CTC / attention decoding, produced by running the REFERENCE's own
classes in the build container (needs /root/reference; not available on the GPU box):

  * CTCPrefixScore.__call__ / initial_state                 (src/models/dicow/decoding.py:8-159)
  * CTCRescorerLogitsProcessor.__call__ / update_state      (src/models/dicow/decoding.py:166-338)

driven the way the greedy branch of _sample drives them (src/models/dicow/generation.py:728-769): for a fixed number of
steps, attention log-probs -> rescorer -> argmax -> finished rows emit pad -> update_state.  The attention scores are a
seeded synthetic function of the step (stored), with -inf entries like the timestamp processor leaves them.

    python tests/golden/make_golden_ctc.py   ->  tests/golden/ctc_joint.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")

from models.dicow.decoding import CTCPrefixScore, CTCRescorerLogitsProcessor  # noqa: E402

# miniature vocabulary with Whisper's ordering: text < eos < specials (sot, lang, task, notimestamps) < timestamps; blank = V
V, EOS, SOT, LANG, TASK, NOTS, TS0, N_TS = 48, 30, 31, 32, 33, 34, 35, 13
BLANK = V
T, B, K, STEPS, W = 14, 4, 9, 11, 0.3


class Tok:
    prefix_tokens = [SOT, LANG, TASK]
    upper_cased_tokens = {3: 20, 5: 21, 7: 22}  # lower id -> upper id share one CTC posterior (decoding.py:183-186)

    def get_vocab(self):
        return {"<|0.00|>": TS0}


def att_scores(rng, step, ids):
    """log-softmaxed attention scores with the masks the timestamp processor would leave: no_timestamps / specials
    are -inf, some steps only allow timestamps or only text"""
    s = torch.from_numpy(rng.normal(size=(B, V)).astype(np.float32)) * 2.0
    s[:, SOT:TS0] = -float("inf")
    if step == 0:
        s[:, :EOS] = -float("inf")  # first token must be a timestamp; DiCoW keeps EOS alive
    if step in (4, 8):
        s[1, TS0:] += 6.0  # a timestamp wins on row 1
    if step >= 6:  # row 2 can only finish
        s[2] = -float("inf")
        s[2, EOS] = 0.0
    return torch.log_softmax(s, dim=-1)


def main():
    rng = np.random.default_rng(11)
    enc_logits = torch.from_numpy(rng.normal(size=(B, T, V + 1)).astype(np.float32)) * 2.5
    enc_logits[..., BLANK] += 1.0
    proc = CTCRescorerLogitsProcessor(enc_logits.clone(), torch.full((B,), T), BLANK, EOS, EOS, SOT, Tok(), 0, W, 1, False,
                                      ctc_tokens_to_score=K)
    ids = torch.tensor([[SOT, LANG, TASK]] * B)
    unfinished = torch.ones(B, dtype=torch.long)
    out = {"enc_logits": enc_logits.numpy(), "meta": np.array([V, EOS, SOT, BLANK, TS0, T, B, K, STEPS]), "ctc_weight": W,
           "upper_lo": np.array(list(Tok.upper_cased_tokens.keys())), "upper_up": np.array(list(Tok.upper_cased_tokens.values()))}
    for step in range(STEPS):
        s = att_scores(rng, step, ids)
        nxt = proc(ids, s.clone())
        tok = torch.argmax(nxt, dim=-1)
        tok = tok * unfinished + EOS * (1 - unfinished)
        proc.update_state(tok, torch.arange(B))
        out[f"att_{step}"] = s.numpy()
        out[f"next_{step}"] = nxt.numpy()
        out[f"tok_{step}"] = tok.numpy()
        out[f"score_prev_{step}"] = proc.ctc_score_prev[:, 0].numpy().copy()
        out[f"state_prev_{step}"] = proc.ctc_state_prev.numpy().copy()
        ids = torch.cat([ids, tok[:, None]], dim=1)
        unfinished = unfinished & (tok != EOS).long()
    out["ids"] = ids.numpy()
    # a direct known-answer case of the prefix scorer alone: two hypotheses with different prefix lengths in one call
    x = torch.log_softmax(enc_logits[:2], dim=-1)
    sc = CTCPrefixScore(x, BLANK, EOS)
    r0, _ = sc.initial_state()
    cs = torch.tensor([[1, 2, 3, EOS, 9], [4, 1, EOS, 6, BLANK]])
    psi0, st0 = sc(torch.tensor([[BLANK], [BLANK]]), cs, torch.tensor([0, 0]), torch.tensor([True, True]), r0)
    out["kat_cs"], out["kat_psi0"], out["kat_r0"] = cs.numpy(), psi0.numpy().copy(), st0.numpy().copy()
    r1 = torch.stack([st0[0, :, :, 1], st0[1, :, :, 0]])  # hypothesis 0 took label 2, hypothesis 1 took label 4
    sc2 = CTCPrefixScore(x, BLANK, EOS)
    y = torch.tensor([[BLANK, 2, 2], [BLANK, 4, 4]])
    psi1, st1 = sc2(y, cs, torch.tensor([2, 1]), torch.tensor([True, True]), r1)
    out["kat_r1_in"], out["kat_psi1"], out["kat_r1"] = r1.numpy().copy(), psi1.numpy().copy(), st1.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "ctc_joint.npz"), **out)
    print("ids:", ids.tolist())
    print("saved", os.path.join(HERE, "ctc_joint.npz"))


if __name__ == "__main__":
    main()
