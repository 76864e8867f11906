"""Golden run of beam search bookkeeping, produced by the helper methods the REFERENCE's `_beam_search` override calls
(src/models/dicow/generation.py:1012-1105: self._get_top_k_continuations, self._get_running_beams_for_next_iteration,
self._update_finished_beams, self._check_early_stop_heuristic, self._beam_search_has_unfinished_sequences -- inherited from
third-party transformers.generation.utils.GenerationMixin; reference pin 4.55.0, run here with the installed version),
called unbound on seeded per-step log-probabilities in exactly the order of the reference's loop.

    python tests/golden/make_golden_beam.py   ->  tests/golden/beam_search.npz
"""
from __future__ import annotations

import os

import numpy as np
import torch
import transformers
from transformers.generation.utils import GenerationMixin as G

HERE = os.path.dirname(os.path.abspath(__file__))
V, EOS, U, K, P, MAXLEN = 37, 5, 3, 4, 3, 12


def run(length_penalty, early_stopping, seed):
    rng = np.random.default_rng(seed)
    dev = "cpu"
    beams_to_keep = 2 * K
    top_mask = torch.cat((torch.ones(K, dtype=torch.bool), torch.zeros(beams_to_keep - K, dtype=torch.bool)))
    running_sequences = torch.full((U, K, MAXLEN), EOS, dtype=torch.int64)
    running_sequences[:, :, :P] = torch.tensor([9, 10, 11])
    sequences = running_sequences.clone()
    running_beam_scores = torch.zeros((U, K))
    running_beam_scores[:, 1:] = -1e9
    beam_scores = torch.full((U, K), -1e9)
    is_sent_finished = torch.zeros((U, K), dtype=torch.bool)
    unsat = torch.ones((U, 1), dtype=torch.bool)
    hits = torch.zeros((U, K), dtype=torch.bool)
    running_beam_indices = torch.full((U, K, MAXLEN - P), -1, dtype=torch.int32)
    beam_indices = running_beam_indices.clone()
    cur_len = P
    out = {}
    step = 0
    while True:
        lp = torch.from_numpy(rng.normal(size=(U * K, V)).astype(np.float32)) * 2.0
        lp[:, EOS] += float(step) * 0.9 - 2.0          # EOS becomes likely as the sequences grow
        lp = torch.log_softmax(lp, dim=-1)
        out[f"lp_{step}"] = lp.numpy()
        acc = (lp.view(U, K, V) + running_beam_scores[:, :, None]).reshape(U, K * V)
        topk_lp, topk_seq, topk_bi = _topk(acc, running_sequences, running_beam_indices, cur_len, beams_to_keep)
        flat = topk_seq[:, :, :cur_len + 1].reshape(U * beams_to_keep, cur_len + 1)
        hits = ((flat[:, -1] == EOS) | (flat.shape[1] >= MAXLEN)).view(U, beams_to_keep)
        running_sequences, running_beam_scores, running_beam_indices = G._get_running_beams_for_next_iteration(
            _Self, topk_log_probs=topk_lp, topk_running_sequences=topk_seq, topk_running_beam_indices=topk_bi,
            next_token_hits_stopping_criteria=hits, num_beams=K)
        sequences, beam_scores, beam_indices, is_sent_finished = G._update_finished_beams(
            _Self, sequences=sequences, topk_running_sequences=topk_seq, beam_scores=beam_scores, topk_log_probs=topk_lp,
            beam_indices=beam_indices, topk_running_beam_indices=topk_bi, is_early_stop_heuristic_unsatisfied=unsat,
            is_sent_finished=is_sent_finished, next_token_hits_stopping_criteria=hits, top_num_beam_mask=top_mask,
            num_beams=K, cur_len=cur_len, decoder_prompt_len=P, length_penalty=length_penalty, early_stopping=early_stopping)
        beam_idx = running_beam_indices[..., cur_len - P].reshape(-1)
        out[f"tok_{step}"] = running_sequences.reshape(U * K, -1)[:, cur_len].numpy().copy()
        out[f"parent_{step}"] = beam_idx.numpy().copy()
        cur_len += 1
        unsat = G._check_early_stop_heuristic(unsat, running_beam_scores, beam_scores, is_sent_finished, cur_len, MAXLEN, P,
                                              early_stopping, length_penalty)
        cont = G._beam_search_has_unfinished_sequences(unsat, is_sent_finished, hits, early_stopping)
        out[f"run_score_{step}"] = running_beam_scores.numpy().copy()
        out[f"fin_score_{step}"] = beam_scores.numpy().copy()
        out[f"fin_flag_{step}"] = is_sent_finished.numpy().copy()
        out[f"unsat_{step}"] = unsat.numpy().copy()
        step += 1
        if not bool(cont):
            break
    out["steps"] = np.array(step)
    out["best"] = sequences[:, 0, :].numpy().copy()
    out["best_score"] = beam_scores[:, 0].numpy().copy()
    return out


class _SelfT:
    _gather_beams = staticmethod(G._gather_beams)


_Self = _SelfT()


def _topk(acc, running_sequences, running_beam_indices, cur_len, beams_to_keep):
    return G._get_top_k_continuations(_Self, accumulated_log_probs=acc, running_sequences=running_sequences,
                                      running_beam_indices=running_beam_indices, cur_len=cur_len, decoder_prompt_len=P,
                                      do_sample=False, beams_to_keep=beams_to_keep, num_beams=K, vocab_size=V, batch_size=U)


if __name__ == "__main__":
    data = {"meta": np.array([V, EOS, U, K, P, MAXLEN]), "transformers_version": np.array(transformers.__version__)}
    cases = {"lp1_noearly": (1.0, False, 1), "lp01_early": (0.1, True, 2), "lp0_never": (0.0, "never", 3)}
    for name, (lpn, early, seed) in cases.items():
        r = run(lpn, early, seed)
        for k, v in r.items():
            data[f"{name}/{k}"] = v
        print(name, "steps", int(r["steps"]), "best", r["best"].tolist(), r["best_score"].tolist())
    np.savez_compressed(os.path.join(HERE, "beam_search.npz"), **data)
