"""CPU oracle (test infrastructure, NOT product code) for beam search with the joint CTC / attention hook -- SURVEY.md 8(f).1.

Restates, one utterance and one candidate at a time (plain Python lists / loops), the beam search the reference runs:
DiCoWGenerationMixin._beam_search (src/models/dicow/generation.py:815-1154), which is HF `GenerationMixin._beam_search`
(third-party `transformers`, reference pin 4.55.0; the installed 5.5.0 has the same helpers) plus one line:
``ctc_rescorer.update_state(next_tokens, beam_idx)`` (generation.py:1087-1088).

Per step (generation.py:992-1107):
  log-probs = log_softmax(logits) -> logits processors (suppress, timestamp rules, CTC rescorer) -> + running beam score
  top-2K continuations of the K x V candidates of an utterance       (_get_top_k_continuations)
  a continuation "hits" when its token is EOS or the length limit     (stopping criteria)
  next running beams = best K continuations that did not hit         (_get_running_beams_for_next_iteration)
  finished set       = best K of {old finished} U {hits among the first K continuations}, scored score / len^penalty
                                                                     (_update_finished_beams)
  early-stop flag    = can the best running beam still beat the worst finished one  (_check_early_stop_heuristic)
The loop ends for the WHOLE batch when no utterance can improve, or every finished slot is full with early_stopping=True,
or every continuation hit (_beam_search_has_unfinished_sequences).  Output: the best finished sequence per utterance.

Pinned: tests/test_oracle_golden.py drives this class and HF's own helper methods (transformers.generation.utils.
GenerationMixin._get_top_k_continuations / _get_running_beams_for_next_iteration / _update_finished_beams /
_check_early_stop_heuristic, called unbound on the same tensors) through identical seeded steps and compares every
intermediate (tests/golden/beam_search.npz holds the stored run of the HF helpers).

Ties: torch.topk leaves the order of equal scores unspecified; this restatement breaks ties by the lower flat index
(beam * V + token), then by the earlier position -- the goldens use continuous random scores, so no ties occur there.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch

NEG = -1.0e9


class BeamSearch:
    """state of one batch: ``seqs[u][k]`` running token lists, ``run_score[u][k]``, finished set per utterance"""

    def __init__(self, prompt: Sequence[Sequence[int]], num_beams: int, *, eos: int, pad: int, max_length: int,
                 length_penalty: float = 1.0, early_stopping=False):
        self.K, self.eos, self.pad, self.max_length = num_beams, eos, pad, max_length
        self.lp, self.early = float(length_penalty), early_stopping
        self.U = len(prompt)
        self.prompt_len = len(prompt[0])
        self.seqs = [[list(p) for _ in range(num_beams)] for p in prompt]
        self.run_score = [[0.0] + [NEG] * (num_beams - 1) for _ in prompt]
        self.fin_seqs: List[List[List[int]]] = [[[] for _ in range(num_beams)] for _ in prompt]
        self.fin_score = [[NEG] * num_beams for _ in prompt]
        self.fin_flag = [[False] * num_beams for _ in prompt]
        self.unsat = [True] * self.U
        self.cur_len = self.prompt_len
        self.last_hits: List[List[bool]] = [[False] * (2 * num_beams) for _ in prompt]
        self.parents: List[List[int]] = [[k for k in range(num_beams)] for _ in prompt]

    def flat_ids(self) -> torch.Tensor:
        return torch.tensor([s for u in self.seqs for s in u], dtype=torch.long)

    def step(self, log_probs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """log_probs [U * K, V] after the logits processors.  Returns (next tokens [U * K], flat parent rows [U * K]) of
        the new running beams -- what the reference hands to ``ctc_rescorer.update_state`` and to the cache reorder."""
        K, V = self.K, log_probs.shape[1]
        f32 = torch.float32
        toks, parents = [], []
        for u in range(self.U):
            acc = (log_probs[u * K:(u + 1) * K].to(f32) + torch.tensor(self.run_score[u], dtype=f32)[:, None]).reshape(-1)
            vals, idx = torch.topk(acc, k=2 * K)
            cand = [(float(vals[j]), int(idx[j]) // V, int(idx[j]) % V) for j in range(2 * K)]
            hits = [(tok == self.eos) or (self.cur_len + 1 >= self.max_length) for _, _, tok in cand]
            # ---- running beams of the next step: best K that did not hit (hits carry -1e9) ----
            runv = torch.tensor([torch.tensor(s, dtype=f32) + torch.tensor(NEG if h else 0.0, dtype=f32) for (s, _, _), h in zip(cand, hits)])
            order = sorted(range(2 * K), key=lambda j: (-float(runv[j]), j))[:K]
            new_seqs = [self.seqs[u][cand[j][1]] + [cand[j][2]] for j in order]
            new_scores = [float(runv[j]) for j in order]
            # ---- finished set ----
            L = float(self.cur_len + 1 - self.prompt_len)
            full = all(self.fin_flag[u]) and (self.early is True)
            merged = [(self.fin_score[u][k], self.fin_flag[u][k], self.fin_seqs[u][k]) for k in range(K)]
            for j, ((s, beam, tok), h) in enumerate(zip(cand, hits)):
                did = h and j < K
                fs = torch.tensor(s, dtype=f32) / torch.tensor(L ** self.lp, dtype=f32)
                fs = fs + torch.tensor(NEG if full else 0.0, dtype=f32)
                fs = fs + torch.tensor(NEG if not self.unsat[u] else 0.0, dtype=f32)
                fs = fs + torch.tensor(NEG if not did else 0.0, dtype=f32)
                merged.append((float(fs), did, self.seqs[u][beam] + [tok]))
            top = sorted(range(len(merged)), key=lambda i: (-merged[i][0], i))[:K]
            self.fin_score[u] = [merged[i][0] for i in top]
            self.fin_flag[u] = [merged[i][1] for i in top]
            self.fin_seqs[u] = [merged[i][2] for i in top]
            self.parents[u] = [cand[j][1] for j in order]
            self.seqs[u], self.run_score[u], self.last_hits[u] = new_seqs, new_scores, hits
            toks += [cand[j][2] for j in order]
            parents += [u * K + cand[j][1] for j in order]
        self.cur_len += 1
        for u in range(self.U):  # _check_early_stop_heuristic
            Lh = (self.max_length - self.prompt_len) if (self.early == "never" and self.lp > 0.0) else (self.cur_len - self.prompt_len)
            best = torch.tensor(self.run_score[u][0], dtype=f32) / torch.tensor(float(Lh) ** self.lp, dtype=f32)
            worst_fin = min(self.fin_score[u])
            worst = [worst_fin if self.fin_flag[u][k] else NEG for k in range(K)]
            self.unsat[u] = self.unsat[u] and any(float(best) > w for w in worst)
        return torch.tensor(toks, dtype=torch.long), torch.tensor(parents, dtype=torch.long)

    def unfinished(self) -> bool:
        improvement = any(self.unsat)
        open_beam = not (all(all(f) for f in self.fin_flag) and (self.early is True))
        valid = not all(all(h) for h in self.last_hits)
        return improvement and open_beam and valid

    def best(self) -> List[List[int]]:
        return [self.fin_seqs[u][0] for u in range(self.U)]


def beam_decode(step_scores: Callable[[torch.Tensor], torch.Tensor], prompt: Sequence[Sequence[int]], num_beams: int, *,
                eos: int, pad: int, max_length: int, length_penalty: float = 1.0, early_stopping=False,
                rescorer=None) -> Tuple[List[List[int]], BeamSearch]:
    """step_scores(ids [U * K, len]) -> processed log-probs [U * K, V] (everything up to, not including, the CTC rescorer);
    ``rescorer`` (oracle.ctc_prefix.JointCtcRescorer over U * K hypotheses) is applied last and told which continuation
    each running beam took (generation.py:1087-1088)."""
    bs = BeamSearch(prompt, num_beams, eos=eos, pad=pad, max_length=max_length, length_penalty=length_penalty,
                    early_stopping=early_stopping)
    while True:
        ids = bs.flat_ids()
        lp = step_scores(ids)
        if rescorer is not None:
            lp = rescorer(ids, lp)
        toks, parents = bs.step(lp)
        if rescorer is not None:
            rescorer.update_state(toks, parents)
        if not bs.unfinished():
            break
    return bs.best(), bs


__all__ = ["BeamSearch", "beam_decode", "NEG"]
