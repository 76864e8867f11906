"""GPU parity of joint CTC / attention decoding (dicow_ctc_joint_step).

  * against the committed golden steps produced by the REFERENCE's own CTCRescorerLogitsProcessor / CTCPrefixScore
    (tests/golden/ctc_joint.npz): chosen tokens identical, carried CTC state / score within 1e-3;
  * against the CPU oracle (oracle/ctc_prefix.py) at Whisper sizes (V = 51 866, T' = 375, top-500 candidates) on seeded
    inputs: candidate sets identical, prefix scores within 1e-3 relative, chosen tokens identical.
Tolerance: fp32 arithmetic on both sides; the differences are libm vs device log / exp (<= 1e-6 per operation, accumulated
over 375 frames), bound 1e-3 relative to the score magnitude as north_star's fp32 bar."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc_prefix as cp

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ops():
    from ts_asr_whisper_b200 import ops as _ops
    return _ops


def test_joint_step_reproduces_reference_golden_steps(ops):
    g = np.load(os.path.join(GOLD, "ctc_joint.npz"))
    V, EOS, SOT, BLANK, TS0, T, B, K, STEPS = [int(v) for v in g["meta"]]
    up = dict(zip(g["upper_lo"].tolist(), g["upper_up"].tolist()))
    st = ops.CtcJointState(torch.from_numpy(g["enc_logits"]).to(DEV).contiguous(), top_k=K, upper_cased=up)
    ids = torch.zeros(B, 3 + STEPS + 1, dtype=torch.int64, device=DEV)
    ids[:, :3] = torch.tensor([SOT, SOT + 1, SOT + 2])
    unf = torch.ones(B, dtype=torch.int32, device=DEV)
    alive = [True] * B
    for step in range(STEPS):
        att = torch.from_numpy(g[f"att_{step}"]).to(DEV).contiguous()  # already log-softmaxed: lse = 0
        ops.ctc_joint_step(st, att, ids, unf, bos=SOT, eos=EOS, pad=EOS, first_timestamp=TS0, prefix_len=3,
                           ctc_weight=float(g["ctc_weight"]), cur_len=3 + step)
        torch.cuda.synchronize()
        tok = ids[:, 3 + step].cpu()
        assert tok.tolist() == g[f"tok_{step}"].tolist(), f"step {step}"
        for b in range(B):
            if alive[b] and int(tok[b]) != EOS:
                np.testing.assert_allclose(st.score_prev[b].item(), g[f"score_prev_{step}"][b], rtol=1e-3, atol=1e-3)
                np.testing.assert_allclose(st.r_prev[b].cpu().numpy(), g[f"state_prev_{step}"][b], rtol=1e-3, atol=1e-3)
            alive[b] = alive[b] and int(tok[b]) != EOS
    assert ids[:, :3 + STEPS].cpu().tolist() == g["ids"].tolist()
    assert unf.cpu().tolist() == [int(a) for a in alive]


def test_joint_step_matches_oracle_at_whisper_sizes(ops):
    """V = 51 866 (+ blank), T' = 375, K = 500, B = 3: three decode steps with different prefix lengths per hypothesis"""
    V, EOS, SOT, TS0, T, B, K, W = 51866, 50257, 50258, 50365, 375, 3, 500, 0.2
    BLANK = V
    rng = np.random.default_rng(2)
    enc_logits = torch.from_numpy(rng.normal(size=(B, T, V + 1)).astype(np.float32)) * 2.0
    enc_logits[..., BLANK] += 6.0  # CTC posteriors are blank-dominated
    resc = cp.JointCtcRescorer(enc_logits, blank=BLANK, eos=EOS, bos=SOT, prefix_len=3, first_timestamp=TS0, ctc_weight=W,
                               top_k=K)
    st = ops.CtcJointState(enc_logits.to(DEV).contiguous(), top_k=K)
    np.testing.assert_allclose(st.logp[0, :3].cpu().numpy(), resc.x[0, :3].numpy(), rtol=1e-5, atol=1e-5)
    prompt = torch.tensor([[SOT, 50259, 50360]] * B)
    steps = 4
    ids_ref = prompt.clone()
    ids = torch.zeros(B, 3 + steps + 1, dtype=torch.int64, device=DEV)
    ids[:, :3] = prompt.to(DEV)
    unf = torch.ones(B, dtype=torch.int32, device=DEV)
    for step in range(steps):
        raw = torch.from_numpy(rng.normal(size=(B, V)).astype(np.float32)) * 3.0
        raw[:, 50258:TS0] = -float("inf")
        if step == 0:
            raw[:, :EOS] = -float("inf")          # first token: timestamps (or EOS) only
            raw[1, TS0 + 7] += 30.0
        if step == 2:
            raw[0, TS0 + 40:TS0 + 60] += 25.0     # hypothesis 0 takes a timestamp: its CTC state must not move
        ref_scores = resc(ids_ref, torch.log_softmax(raw, dim=-1))
        ref_tok = torch.argmax(ref_scores, dim=-1)
        ops.ctc_joint_step(st, raw.to(DEV).contiguous(), ids, unf, bos=SOT, eos=EOS, pad=EOS, first_timestamp=TS0,
                           prefix_len=3, ctc_weight=W, cur_len=3 + step)
        torch.cuda.synchronize()
        cand = st.candidates.cpu()
        psi = st.prefix_scores.cpu()
        lsm = torch.log_softmax(raw, dim=-1)
        for b in range(B):
            rc = resc._cand[b]
            # candidates whose attention score is -inf (masked ids topk had to fill the set with) are chosen among ties
            # in an unspecified order by torch.topk and can never win: compare the scored ids that matter
            live = lambda ids_: sorted(c for c in ids_ if torch.isfinite(lsm[b, c]))  # noqa: E731
            assert live(cand[b].tolist()) == live(rc["cs"]), f"step {step} row {b}: candidate sets differ"
            order = {c: j for j, c in enumerate(rc["cs"])}
            mine = {int(c): float(psi[b, j]) for j, c in enumerate(cand[b])}
            for c in live(rc["cs"]):
                np.testing.assert_allclose(mine[c], float(rc["psi"][order[c]]), rtol=1e-3, atol=1e-2)
        assert ids[:, 3 + step].cpu().tolist() == ref_tok.tolist(), f"step {step}"
        resc.update_state(ref_tok)
        for b in range(B):
            np.testing.assert_allclose(st.score_prev[b].item(), float(resc.score_prev[b]), rtol=1e-3, atol=1e-2)
            np.testing.assert_allclose(st.r_prev[b].cpu().numpy(), resc.r_prev[b].numpy(), rtol=1e-3, atol=1e-2)
        ids_ref = torch.cat([ids_ref, ref_tok[:, None]], dim=1)


@pytest.mark.parametrize("graphs", [False, True])
def test_joint_ctc_greedy_decode_matches_oracle(graphs):
    """greedy_decode_window(ctc=...) on the miniature model: decoder step + suppress / timestamp rules + log-softmax + CTC
    rescoring + argmax + update_state per token (generation.py:728-769 with generation_config.ctc_weight > 0), against the
    same loop built from the oracle's pieces.  Token identity; a divergence is tolerated only where the oracle's own
    top-2 margin of the combined score is below MARGIN (bf16 decoder vs fp32 oracle)."""
    import test_gpu_decoder as tgd
    from oracle import dicow_oracle as orc
    from oracle import synth
    import torch.nn.functional as F
    dm = synth.GOLDEN_MINI
    dmp = synth.Dims(**{**dm.__dict__, "use_enrollments": False, "scb_layers": 0})
    model, p = tgd.build_model(dmp)
    model.use_cuda_graphs = graphs
    feats = torch.from_numpy(synth.make_features("g0", 2, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno("g0", 2, dm.T, "soft", pad_tail=7))
    W, K, NEW = 0.3, 40, 20
    prompt = torch.tensor([[tgd.SOT, tgd.LANG, tgd.TASK]] * 2)
    enc_mod = model.get_encoder()
    hidden = enc_mod(feats.to(DEV), stno_mask=stno.to(DEV)).last_hidden_state
    ctc_logits = model.get_enc_logits(hidden)
    rules = dict(eos=tgd.EOS, pad=tgd.EOS, no_timestamps=tgd.NOTS, ts_begin=tgd.TS_BEGIN, max_initial_timestamp_index=None,
                 timestamp_rules=True, suppress_bitmap=model._suppress_bitmap(tgd.SUPPRESS, torch.device(DEV)))
    ctc = {"logits": ctc_logits, "weight": W, "prefix_len": 3, "bos": tgd.SOT, "top_k": K}
    ids = model.greedy_decode_window(hidden, prompt.to(DEV), 3 + NEW, rules, ctc=ctc)
    ids2 = model.greedy_decode_window(hidden, prompt.to(DEV), 3 + NEW, rules, ctc=ctc)  # buffers / graphs reused
    torch.cuda.synchronize()
    assert torch.equal(ids, ids2)
    ids = ids.cpu()
    # ---- the same loop from the oracle's pieces ----
    with torch.no_grad():
        ref_enc = orc.encoder_forward(p, dmp, feats, stno)
        ref_ctc = orc.ctc_logits(p, dmp, ref_enc)
        err = ((ctc_logits.cpu() - ref_ctc).abs().max() / ref_ctc.abs().max()).item()
        assert err < 2e-2
        resc = cp.JointCtcRescorer(ref_ctc, blank=dmp.vocab, eos=tgd.EOS, bos=tgd.SOT, prefix_len=3,
                                   first_timestamp=tgd.TS_BEGIN, ctc_weight=W, top_k=K)
        ref_ids = prompt.clone()
        unfinished = torch.ones(2, dtype=torch.long)
        sup = torch.tensor(tgd.SUPPRESS)
        n_same, diverged = 0, [False, False]
        for step in range(NEW):
            hid = orc.decoder_forward(p, dmp, ref_ids, ref_enc)
            logits = F.linear(hid[:, -1], p["proj_out.weight"]).float()
            logits[:, sup] = -float("inf")
            logits = orc.timestamp_rules(ref_ids, logits, begin_index=3, eos=tgd.EOS, no_timestamps=tgd.NOTS,
                                         ts_begin=tgd.TS_BEGIN)
            comb = resc(ref_ids, torch.log_softmax(logits, dim=-1))
            tok = torch.argmax(comb, dim=-1)
            tok = tok * unfinished + tgd.EOS * (1 - unfinished)
            for b in range(2):
                if diverged[b] or 3 + step >= ids.shape[1]:
                    continue
                got = int(ids[b, 3 + step])
                if got == int(tok[b]):
                    n_same += 1
                    continue
                margin = float(comb[b, int(tok[b])] - comb[b, got])
                assert margin < tgd.MARGIN, f"row {b} step {step}: token {got} vs {int(tok[b])}, margin {margin:.3f}"
                diverged[b] = True  # after a tolerated near-tie flip the continuations legitimately differ
            resc.update_state(tok)
            ref_ids = torch.cat([ref_ids, tok[:, None]], dim=1)
            unfinished = unfinished & (tok != tgd.EOS).long()
            if int(unfinished.max()) == 0:
                break
    print("joint greedy ids (cuda):", ids.tolist(), "\n                (oracle):", ref_ids.tolist())
    assert n_same >= 12, "too few compared tokens for the check to mean anything"
