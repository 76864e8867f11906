"""CPU oracle (test infrastructure, NOT product code) for joint CTC / attention decoding -- SURVEY.md section 8(f).1.

Restates, hypothesis by hypothesis and frame by frame (a Python loop over hypotheses and over the CTC frames, torch
vectors over the candidates of one hypothesis), what the reference computes with tensor code batched over hypotheses:

  * ``prefix_scores``   -- CTCPrefixScore.__call__            (src/models/dicow/decoding.py:8-159; ESPnet's vectorised form of
                           Watanabe et al., "Hybrid CTC/attention architecture for end-to-end speech recognition", Alg. 2)
  * ``JointCtcRescorer`` -- CTCRescorerLogitsProcessor         (src/models/dicow/decoding.py:166-338): prefix bookkeeping on
                           the token ids, top-k candidate choice, score combination, state selection (``update_state``)

Pinned: tests/test_oracle_golden.py replays tests/golden/ctc_joint.npz, which tests/golden/make_golden_ctc.py produced by
running the REFERENCE's own two classes (imported from /root/reference/src) on seeded synthetic inputs.

Reference behaviours kept on purpose (they change numbers):
  * LOGZERO is -1e10, not -inf, and enters logaddexp / logsumexp as a number;
  * ``decoded_len`` counts generated ids ``<= first_timestamp`` (the ``<|0.00|>`` token itself is counted, decoding.py:277),
    while "is a timestamp" is ``>= first_timestamp`` (decoding.py:279) and "is not a timestamp" is ``< first_timestamp``;
  * a trailing timestamp is replaced by ``ids[count_of_non_timestamps - 1]`` -- an index by COUNT (decoding.py:281-285),
    which is the last non-timestamp token only when no timestamp precedes it;
  * the frame loop of the forward variables starts at ``min over the scored hypotheses of max(decoded_len, 1)``
    (decoding.py:98-107): a hypothesis with a longer prefix than another one in the same call gets forward variables for
    frames its prefix cannot have reached (the per-frame mask at decoding.py:105-106 never fires because the loop starts at
    the minimum).  The prefix scores (log psi) are masked correctly (decoding.py:91-95) and do not depend on it; the STATES
    handed to the next step do.
  * timestamp ids get the row maximum of the CTC scores (decoding.py:325), candidates outside the top-k get LOGZERO.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

LOGZERO = -1e10


def _lae(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return torch.logaddexp(a, b)


def initial_state(x: torch.Tensor, blank: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """x: log-probs [T, V1] of one utterance -> (r [T, 2] with r[:, 1] = cumsum of the blank log-probs, score 0);
    decoding.py:37-44"""
    T = x.shape[0]
    r = torch.full((T, 2), LOGZERO, dtype=x.dtype)
    r[:, 1] = torch.cumsum(x[:, blank], dim=0)
    return r, torch.zeros((), dtype=x.dtype)


def prefix_scores(x: torch.Tensor, cs: Sequence[int], last: int, decoded_len: int, r_prev: torch.Tensor, blank: int,
                  eos: int, loop_start: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """One hypothesis.  x [T, V1] log-probs, cs candidate ids, ``last`` the last label of the prefix, ``decoded_len`` its
    length, r_prev [T, 2] the forward variables of the prefix, ``loop_start`` the first frame of the forward recursion
    (see the module docstring).  Returns (log_psi [C], r [T, 2, C])."""
    T = x.shape[0]
    C = len(cs)
    dt = x.dtype
    cs_t = torch.tensor(list(cs), dtype=torch.long)
    xs = x[:, cs_t]                                                  # [T, C] posteriors of the candidates
    r = torch.full((T, 2, C), LOGZERO, dtype=dt)
    r_sum = _lae(r_prev[:, 0], r_prev[:, 1])                         # log(r^n(g) + r^b(g)) of the prefix g
    # log phi: paths of the prefix that may be extended by c at the next frame -- all of them, except for c == last label
    # where only the paths ending in blank count (a repeated label needs a blank in between)
    phi = r_sum[:, None].repeat(1, C)
    if decoded_len > 0:
        phi[:, cs_t == last] = r_prev[:, 1:2]
    if decoded_len == 0:
        r[0, 0] = xs[0]
    start = max(decoded_len, 1)
    psi = r[start - 1, 0].clone()
    if T > 1:
        terms = phi[:-1] + xs[1:]                                    # new label starts at frame t = 1 .. T-1
        frames = torch.arange(1, T)[:, None]
        terms = torch.where(frames >= decoded_len, terms, torch.full_like(terms, LOGZERO))
        psi = _lae(psi, torch.logsumexp(terms, dim=0))
    for t in range(loop_start, T):                                   # frame by frame (decoding.py:100-104)
        r[t, 0] = _lae(r[t - 1, 0], phi[t - 1]) + xs[t]
        r[t, 1] = _lae(r[t - 1, 0], r[t - 1, 1]) + x[t, blank]
    psi[cs_t == eos] = r_sum[T - 1]
    if eos != blank:
        psi[cs_t == blank] = LOGZERO
    return psi, r


class JointCtcRescorer:
    """CTCRescorerLogitsProcessor restated for a batch of hypotheses, keeping the forward variables of the scored candidates
    only ([C] per hypothesis instead of the reference's dense [V] table)."""

    def __init__(self, enc_logits: torch.Tensor, *, blank: int, eos: int, bos: int, prefix_len: int, first_timestamp: int,
                 ctc_weight: float, top_k: int = 500, upper_cased: Optional[dict] = None):
        logp = torch.log_softmax(enc_logits, dim=-1)
        if upper_cased:  # decoding.py:183-186: upper-cased variants share the lower-cased token's CTC posterior
            lo = torch.tensor(list(upper_cased.keys()))
            up = torch.tensor(list(upper_cased.values()))
            logp[..., up] = logp[..., lo]
        self.x = logp
        self.blank, self.eos, self.bos, self.prefix_len = blank, eos, bos, prefix_len
        self.first_ts, self.w, self.k = first_timestamp, ctc_weight, top_k
        st = [initial_state(self.x[b], blank) for b in range(self.x.shape[0])]
        self.r_prev = [s[0] for s in st]
        self.score_prev = [s[1] for s in st]
        self._cand: List[Optional[dict]] = [None] * self.x.shape[0]

    def _prefix(self, row: torch.Tensor) -> Tuple[List[int], int, int]:
        """token ids of one hypothesis -> (labels with the sos slot, decoded_len, last label); decoding.py:263-286"""
        ids = [int(v) for v in row]
        if ids[0] != self.bos:
            ids = ids[ids.index(self.bos):]
        if self.prefix_len > 1:
            ids = ids[self.prefix_len - 1:]
        ids[0] = self.blank
        decoded_len = sum(1 for v in ids if v <= self.first_ts and v != self.blank)
        if ids[-1] >= self.first_ts and ids[-1] != self.blank:
            n_text = sum(1 for v in ids if v < self.first_ts or v == self.blank)
            ids[-1] = ids[n_text - 1]
        return ids, decoded_len, ids[-1]

    def __call__(self, input_ids: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
        """scores: attention log-probs [B, V] after the other processors -> (1 - w) scores + w (ctc - ctc_prev)"""
        B, V = scores.shape
        info = [self._prefix(input_ids[b]) for b in range(B)]
        todo = [b for b in range(B) if info[b][2] != self.eos]
        loop_start = min([max(info[b][1], 1) for b in todo], default=1)
        out = torch.empty_like(scores)
        for b in range(B):
            ctc = torch.full((V,), LOGZERO, dtype=scores.dtype)
            self._cand[b] = None
            if b in todo:
                cs = torch.topk(scores[b, :self.first_ts], k=self.k).indices.tolist()
                if self.eos not in cs:
                    cs[self.k - 1] = self.eos
                _, dl, last = info[b]
                psi, r = prefix_scores(self.x[b], cs, last, dl, self.r_prev[b], self.blank, self.eos, loop_start)
                ctc[torch.tensor(cs)] = psi.to(scores.dtype)
                self._cand[b] = {"cs": cs, "psi": psi, "r": r}
            ctc[self.first_ts:] = ctc.max()
            out[b] = (1 - self.w) * scores[b] + self.w * (ctc - self.score_prev[b])
        return out

    def update_state(self, best_ids: torch.Tensor, beam_idx: Optional[Sequence[int]] = None) -> None:
        """decoding.py:253-260: a chosen text token moves the hypothesis to that candidate's forward variables / score; a
        timestamp keeps the parent's.  ``beam_idx[b]`` = the parent hypothesis of new hypothesis b."""
        B = len(best_ids)
        beam_idx = list(range(B)) if beam_idx is None else [int(v) for v in beam_idx]
        new_r, new_s = [], []
        for b in range(B):
            parent, tok = beam_idx[b], int(best_ids[b])
            cand = self._cand[parent]
            if tok < self.first_ts:
                if cand is not None and tok in cand["cs"]:
                    j = cand["cs"].index(tok)
                    new_r.append(cand["r"][:, :, j].clone())
                    new_s.append(cand["psi"][j].clone())
                else:  # the reference reads its dense table: LOGZERO score; the state slot holds stale memory there
                    new_r.append(torch.full_like(self.r_prev[parent], LOGZERO))
                    new_s.append(torch.tensor(LOGZERO, dtype=self.x.dtype))
            else:
                new_r.append(self.r_prev[parent])
                new_s.append(self.score_prev[parent])
        self.r_prev, self.score_prev = new_r, new_s


def greedy_joint_decode(rescorer: JointCtcRescorer, att_scores_fn, prompt: torch.Tensor, steps: int, *, eos: int, pad: int
                        ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Drive the rescorer the way the greedy branch of _sample does (src/models/dicow/generation.py:728-769):
    scores = att_scores_fn(ids) [B, V] (already log-softmaxed) -> rescorer -> argmax -> pad finished rows -> update_state."""
    ids = prompt.clone()
    unfinished = torch.ones(ids.shape[0], dtype=torch.long)
    history = []
    for _ in range(steps):
        nxt_scores = rescorer(ids, att_scores_fn(ids))
        history.append(nxt_scores)
        tok = torch.argmax(nxt_scores, dim=-1)
        tok = tok * unfinished + pad * (1 - unfinished)
        rescorer.update_state(tok)
        ids = torch.cat([ids, tok[:, None]], dim=1)
        unfinished = unfinished & (tok != eos).long()
        if int(unfinished.max()) == 0:
            break
    return ids, history


__all__ = ["LOGZERO", "initial_state", "prefix_scores", "JointCtcRescorer", "greedy_joint_decode"]
