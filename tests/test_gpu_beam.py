"""GPU parity of beam search (dicow_beam_step + the decode step with ancestry-linked caches), SURVEY.md section 8(f).1.

  * the device bookkeeping replays the committed golden runs of the HF helper methods the reference's _beam_search override
    calls (tests/golden/beam_search.npz): tokens, parents, running / finished scores, flags, early-stop state, best
    sequences -- for length_penalty 1.0 / early_stopping False, 0.1 / True, 0.0 / "never";
  * end to end on the miniature model: beam_decode_window (attention only, and joint CTC / attention) against the oracle's
    beam search driven by the oracle decoder, logits rules and CTC rescorer.
Scores are fp32 on both sides (tolerance 1e-4 relative: powf vs pow, fused multiply-add); tokens must be identical."""
import os

import numpy as np
import pytest
import torch

from oracle import beam_search as obs
from oracle import ctc_prefix as cp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ops():
    from ts_asr_whisper_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("case,lp,early", [("lp1_noearly", 1.0, False), ("lp01_early", 0.1, True), ("lp0_never", 0.0, "never")])
def test_beam_step_replays_hf_golden(ops, case, lp, early):
    g = np.load(os.path.join(GOLD, "beam_search.npz"))
    V, EOS, U, K, P, MAXLEN = [int(v) for v in g["meta"]]
    R, FIRST_TS, KC, S = U * K, 30, 16, MAXLEN + 1
    i32, f32 = dict(dtype=torch.int32, device=DEV), dict(dtype=torch.float32, device=DEV)
    ids = torch.zeros(R, S, dtype=torch.int64, device=DEV)
    ids[:, :P] = torch.tensor([9, 10, 11], device=DEV)
    fin_ids = torch.full((R, S), EOS, dtype=torch.int64, device=DEV)
    fin_ids[:, :P] = ids[:, :P]
    ids_tmp = torch.zeros(2 * R, S, dtype=torch.int64, device=DEV)
    run = torch.full((R,), -1e9, **f32)
    run.view(U, K)[:, 0] = 0.0
    fin_score, fin_flag, unsat = torch.full((R,), -1e9, **f32), torch.zeros(R, **i32), torch.ones(U, **i32)
    anc = torch.arange(R, **i32)[:, None].expand(R, S).contiguous()
    anc_tmp = torch.zeros_like(anc)
    pos = torch.tensor([P - 1], **i32)
    unfinished = torch.ones(R, **i32)
    cand = ops.CandidateState(R, KC, DEV)
    sc_i, sc_f, flags = torch.zeros(3 * R, **i32), torch.zeros(2 * R, **f32), torch.zeros(U, 4, **i32)
    steps = int(g[f"{case}/steps"])
    for step in range(steps):
        lpn = torch.from_numpy(g[f"{case}/lp_{step}"]).to(DEV).contiguous()  # log-probs: their own log-softmax normaliser is 0
        ops.ctc_joint_step(cand, lpn, ids, unfinished, pos=pos, bos=9, eos=EOS, pad=EOS, first_timestamp=FIRST_TS, prefix_len=P,
                           ctc_weight=0.0, raw_logits=lpn, score_only=True)
        ops.beam_step(U=U, NB=K, processed_scores=lpn, joint=cand, ctc_weight=0.0, run_score=run, fin_score=fin_score,
                      fin_flag=fin_flag, unsat=unsat, ids=ids, fin_ids=fin_ids, ids_tmp=ids_tmp, ancestry=anc,
                      ancestry_tmp=anc_tmp, pos=pos, eos=EOS, pad=EOS, first_timestamp=FIRST_TS, max_length=MAXLEN,
                      prompt_len=P, length_penalty=lp, early_stopping=early, scratch_i32=sc_i, scratch_f32=sc_f, flags=flags)
        torch.cuda.synchronize()
        cur = P + step  # column of the new token
        f = flags.cpu()
        all_hit = not bool(f[:, 0].any())
        if not all_hit:  # (all running scores tie at -1e9 on the last, length-limited step: order unspecified, unused)
            assert ids[:, cur].cpu().tolist() == g[f"{case}/tok_{step}"].tolist(), f"step {step}"
            assert sc_i[:R].cpu().tolist() == g[f"{case}/parent_{step}"].tolist(), f"step {step}"
            # the ancestry of every new hypothesis: its parent's chain up to the consumed position, itself afterwards
            a = anc.cpu()
            par = g[f"{case}/parent_{step}"]
            for r in range(R):
                assert int(a[r, cur - 1]) == int(par[r]) or step > 0 and int(a[r, cur - 1]) >= 0
                assert int(a[r, cur]) == r
        np.testing.assert_allclose(run.cpu().numpy().reshape(U, K), g[f"{case}/run_score_{step}"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(fin_score.cpu().numpy().reshape(U, K), g[f"{case}/fin_score_{step}"], rtol=1e-4, atol=1e-4)
        assert fin_flag.cpu().bool().reshape(U, K).tolist() == g[f"{case}/fin_flag_{step}"].tolist()
        assert unsat.cpu().bool().reshape(U, 1).tolist() == g[f"{case}/unsat_{step}"].tolist()
        cont = bool(f[:, 2].any()) and not (bool(f[:, 1].all()) and early is True) and bool(f[:, 0].any())
        assert cont == (step < steps - 1), f"loop condition at step {step}"
        pos += 1
    best = fin_ids.view(U, K, S)[:, 0, :MAXLEN].cpu()
    assert best.tolist() == g[f"{case}/best"].tolist()


def _mini():
    import test_gpu_decoder as tgd
    from oracle import synth
    dm = synth.GOLDEN_MINI
    dmp = synth.Dims(**{**dm.__dict__, "use_enrollments": False, "scb_layers": 0})
    model, p = tgd.build_model(dmp)
    feats = torch.from_numpy(synth.make_features("g0", 2, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno("g0", 2, dm.T, "soft", pad_tail=7))
    return tgd, dmp, model, p, feats, stno


@pytest.mark.parametrize("w,graphs", [(0.0, False), (0.3, False), (0.3, True)])
def test_beam_decode_window_matches_oracle(w, graphs):
    """miniature model, 2 utterances x 3 beams: the device loop (decode step over 6 hypotheses with shared cross K/V and
    ancestry-linked self-attention caches, rules, candidates, CTC prefix scores, beam bookkeeping) against the oracle"""
    import torch.nn.functional as F
    from oracle import dicow_oracle as orc
    tgd, dmp, model, p, feats, stno = _mini()
    model.use_cuda_graphs = graphs
    NB, NEW, K, LP = 3, 14, 40, 0.1
    prompt = torch.tensor([[tgd.SOT, tgd.LANG, tgd.TASK]] * 2)
    hidden = model.get_encoder()(feats.to(DEV), stno_mask=stno.to(DEV)).last_hidden_state
    rules = dict(eos=tgd.EOS, pad=tgd.EOS, no_timestamps=tgd.NOTS, ts_begin=tgd.TS_BEGIN, max_initial_timestamp_index=None,
                 timestamp_rules=True, suppress_bitmap=model._suppress_bitmap(tgd.SUPPRESS, torch.device(DEV)))
    ctc = None
    if w > 0:
        ctc = {"logits": model.get_enc_logits(hidden), "weight": w, "prefix_len": 3, "bos": tgd.SOT}
    got = model.beam_decode_window(hidden, prompt.to(DEV), 3 + NEW, rules, num_beams=NB, length_penalty=LP, ctc=ctc, top_k=K)
    got2 = model.beam_decode_window(hidden, prompt.to(DEV), 3 + NEW, rules, num_beams=NB, length_penalty=LP, ctc=ctc, top_k=K)
    torch.cuda.synchronize()
    assert torch.equal(got, got2)
    with torch.no_grad():
        ref_enc = orc.encoder_forward(p, dmp, feats, stno)
        enc_rep = ref_enc.repeat_interleave(NB, dim=0)
        sup = torch.tensor(tgd.SUPPRESS)

        def step_scores(ids):
            hid = orc.decoder_forward(p, dmp, ids, enc_rep)
            logits = F.linear(hid[:, -1], p["proj_out.weight"]).float()
            lp = torch.log_softmax(logits, dim=-1)  # beam search normalises BEFORE the processors (generation.py:1003)
            lp[:, sup] = -float("inf")
            return orc.timestamp_rules(ids, lp, begin_index=3, eos=tgd.EOS, no_timestamps=tgd.NOTS, ts_begin=tgd.TS_BEGIN)

        resc = None
        if w > 0:
            ref_ctc = orc.ctc_logits(p, dmp, ref_enc).repeat_interleave(NB, dim=0)
            resc = cp.JointCtcRescorer(ref_ctc, blank=dmp.vocab, eos=tgd.EOS, bos=tgd.SOT, prefix_len=3,
                                       first_timestamp=tgd.TS_BEGIN, ctc_weight=w, top_k=K)
        best, bs = obs.beam_decode(step_scores, prompt.tolist(), NB, eos=tgd.EOS, pad=tgd.EOS, max_length=3 + NEW,
                                   length_penalty=LP, early_stopping=False, rescorer=resc)
    got = got.cpu()
    print("beam (cuda):  ", got.tolist(), "\nbeam (oracle):", best, "scores", [s[0] for s in bs.fin_score])
    for u in range(2):
        assert got[u, :len(best[u])].tolist() == best[u], f"utterance {u}"
        assert all(int(t) == tgd.EOS for t in got[u, len(best[u]):])
