"""Pin the CPU oracle (oracle/dicow_oracle.py) to outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference's own modules).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import dicow_oracle as orc
from oracle import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
EOS, SOT, LANG, TASK, NOTS, TS_BEGIN, N_TS = 257, 258, 259, 260, 261, 262, 38
SUPPRESS = [1, 2, 7, 8, 9, 10, 14, 25, 258, 259, 260]


@pytest.fixture(scope="module")
def mini():
    g = np.load(os.path.join(GOLD, "mini_model.npz"))
    dm = synth.GOLDEN_MINI
    p = orc.to_torch(synth.make_params(dm))
    B = 2
    feats = torch.from_numpy(synth.make_features("g0", B, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno("g0", B, dm.T, "soft", pad_tail=7))
    enr = {"input_features": torch.from_numpy(synth.make_features("g0e", B, dm.n_mels, 2 * dm.T)),
           "stno_mask": torch.from_numpy(synth.make_stno("g0e", B, dm.T, "hard"))}
    dm_plain = synth.Dims(**{**dm.__dict__, "use_enrollments": False, "scb_layers": 0})
    return g, dm, dm_plain, p, feats, stno, enr


def _close(a, b, tol=2e-5):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else a
    err = np.abs(a - b).max()
    assert err <= tol * max(1.0, np.abs(b).max()), f"max abs err {err}"


def test_encoder_se_dicow(mini):
    g, dm, _, p, feats, stno, enr = mini
    with torch.no_grad():
        _close(orc.encoder_forward(p, dm, feats, stno, enr), g["enc_se"])
        _close(orc.encoder_forward(p, dm, feats, stno, enr, return_logits=True), g["ctc_logits_se"])


def test_encoder_plain(mini):
    g, _, dmp, p, feats, stno, _ = mini
    with torch.no_grad():
        _close(orc.encoder_forward(p, dmp, feats, stno), g["enc_plain"])


def test_stno_rows_sum_to_one(mini):
    _, dm, _, _, _, stno, enr = mini
    assert torch.allclose(stno.sum(1), torch.ones_like(stno.sum(1)), atol=1e-6)
    assert torch.allclose(enr["stno_mask"].sum(1), torch.ones(2, dm.T))


def test_forward_losses(mini):
    g, _, dmp, p, feats, stno, _ = mini
    labels = torch.from_numpy(g["labels"])
    upp = torch.from_numpy(g["upp_labels"])
    assert np.array_equal(g["labels"], synth.make_labels("g0", 2, 12, dmp.vocab, EOS, TS_BEGIN, prefix=(LANG, TASK)))
    with torch.no_grad():
        loss, logits, _ = orc.model_forward(p, dmp, feats, stno, labels, upp, ctc_prefix_tokens=(SOT, LANG, TASK))
        _close(logits, g["fwd_logits"], 5e-5)
        _close(loss, g["fwd_hard_loss"], 2e-5)
        loss2, _, _ = orc.model_forward(p, dmp, feats, stno, labels, upp, ctc_prefix_tokens=(SOT, LANG, TASK),
                                        ts_begin=TS_BEGIN, n_ts=N_TS)
        _close(loss2, g["fwd_soft_loss"], 2e-5)


def test_greedy_decode_token_identical(mini):
    g, _, dmp, p, feats, stno, _ = mini
    with torch.no_grad():
        enc = orc.encoder_forward(p, dmp, feats, stno)
        prompt = torch.tensor([[SOT, LANG, TASK]] * 2)
        ids, lgs = orc.greedy_decode(p, dmp, enc, prompt, 24, suppress=SUPPRESS, no_timestamps=NOTS, ts_begin=TS_BEGIN,
                                     return_logits=True)
    _close(lgs[0], g["greedy_first_logits"], 5e-5)
    assert ids.tolist() == g["greedy_ids"].tolist()


def test_fddt_identity_at_init():
    """Known-answer identity: per-layer FDDT with all weights 1, biases 0 is the identity when the
    STNO rows sum to 1 (reference init: src/models/dicow/encoder.py:49-61)."""
    x = torch.from_numpy(synth.gaussish("kat/x", (2, 9, 16)))
    stno = torch.from_numpy(synth.make_stno("kat", 2, 9, "soft"))
    y = orc.fddt(x, stno, torch.ones(4, 16), torch.zeros(4, 16))
    assert torch.allclose(x, y, atol=1e-6)
    # initial_fddt at init: x * (v m_S + m_T + v m_N + m_O), v = non_target_fddt_value (encoder.py:62-73)
    v = 0.5
    w = torch.tensor([v, 1.0, v, 1.0])[:, None].expand(4, 16)
    y = orc.fddt(x, stno, w, torch.zeros(4, 16))
    scale = v * stno[:, 0] + stno[:, 1] + v * stno[:, 2] + stno[:, 3]
    assert torch.allclose(y, x * scale[..., None], atol=1e-6)


@pytest.mark.parametrize("n_mels", [80, 128])
def test_log_mel(n_mels):
    g = np.load(os.path.join(GOLD, "mel.npz"))
    _close(orc.mel_filterbank(n_mels), g[f"filters{n_mels}"], 1e-6)
    wav = synth.make_audio(f"mel{n_mels}", 41777)
    feat, mask = orc.log_mel(wav, n_mels, chunk_samples=32000)
    assert feat.shape == g[f"feat{n_mels}"].shape
    _close(feat, g[f"feat{n_mels}"], 2e-5)
    assert np.array_equal(mask, g[f"mask{n_mels}"])


@pytest.mark.parametrize("case", range(4))
def test_log_mel_equals_installed_whisper_extractor_on_random_lengths(case):
    """the log-mel oracle against the installed HF WhisperFeatureExtractor called exactly as the reference does
    (src/data/local_datasets.py:208-214), 30 s windows: sub-window, exact, and multi-window recordings (shared floor)"""
    from transformers import WhisperFeatureExtractor
    rng = np.random.default_rng(60 + case)
    n_mels = 80 if case % 2 else 128
    n = int([rng.integers(4000, 400000), 480000, rng.integers(480001, 900000), rng.integers(960001, 1100000)][case])
    wav = (0.1 * rng.standard_normal(n)).astype(np.float32)
    wav[: n // 3] *= 0.01  # a quiet stretch: exercises the (max - 8) floor
    fe = WhisperFeatureExtractor(feature_size=n_mels)
    f = fe(wav, return_tensors="pt", sampling_rate=16000, return_attention_mask=True, truncation=False, padding="longest",
           pad_to_multiple_of=fe.n_samples)
    feat, mask = orc.log_mel(wav, n_mels)
    assert feat.shape == tuple(f.input_features[0].shape)
    _close(feat, f.input_features[0].numpy(), 5e-5)
    assert np.array_equal(mask, f.attention_mask[0].numpy())


# ---- joint CTC / attention decoding: oracle/ctc_prefix.py vs the reference's own classes --------
def _ctc_golden():
    return np.load(os.path.join(GOLD, "ctc_joint.npz"))


def test_ctc_prefix_scorer_known_answers():
    """CTCPrefixScore.__call__ (decoding.py:121-159) on two hypotheses per call, first with empty prefixes, then with
    prefixes of different lengths (the forward recursion starts at the SHORTER prefix for both: module docstring)"""
    from oracle import ctc_prefix as cp
    g = _ctc_golden()
    V, EOS, SOT, BLANK, TS0, T, B, K, STEPS = [int(v) for v in g["meta"]]
    x = torch.log_softmax(torch.from_numpy(g["enc_logits"][:2]), dim=-1)
    cs = g["kat_cs"]
    for b in range(2):
        r0, _ = cp.initial_state(x[b], BLANK)
        psi, r = cp.prefix_scores(x[b], cs[b].tolist(), BLANK, 0, r0, BLANK, EOS, loop_start=1)
        np.testing.assert_allclose(psi.numpy(), g["kat_psi0"][b], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(r.numpy(), g["kat_r0"][b], rtol=2e-5, atol=2e-5)
    lens, lasts = [2, 1], [2, 4]
    for b in range(2):
        psi, r = cp.prefix_scores(x[b], cs[b].tolist(), lasts[b], lens[b], torch.from_numpy(g["kat_r1_in"][b]), BLANK, EOS,
                                  loop_start=min(lens))
        np.testing.assert_allclose(psi.numpy(), g["kat_psi1"][b], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(r.numpy(), g["kat_r1"][b], rtol=2e-5, atol=2e-5)


def test_joint_ctc_rescorer_matches_reference_steps():
    """CTCRescorerLogitsProcessor driven like the greedy loop (generation.py:728-769): per step the combined scores, the
    chosen tokens and the carried CTC state / score of every unfinished hypothesis"""
    from oracle import ctc_prefix as cp
    g = _ctc_golden()
    V, EOS, SOT, BLANK, TS0, T, B, K, STEPS = [int(v) for v in g["meta"]]
    up = dict(zip(g["upper_lo"].tolist(), g["upper_up"].tolist()))
    resc = cp.JointCtcRescorer(torch.from_numpy(g["enc_logits"]), blank=BLANK, eos=EOS, bos=SOT, prefix_len=3,
                               first_timestamp=TS0, ctc_weight=float(g["ctc_weight"]), top_k=K, upper_cased=up)
    ids = torch.tensor([[SOT, SOT + 1, SOT + 2]] * B)
    unfinished = torch.ones(B, dtype=torch.long)
    n_text = 0
    for step in range(STEPS):
        nxt = resc(ids, torch.from_numpy(g[f"att_{step}"]))
        ref = torch.from_numpy(g[f"next_{step}"])
        live = ref > -1e8  # candidates and timestamps; everything else carries w * LOGZERO
        assert torch.equal(live, nxt > -1e8)
        np.testing.assert_allclose(nxt[live].numpy(), ref[live].numpy(), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(nxt[~live & torch.isfinite(ref)].numpy(), ref[~live & torch.isfinite(ref)].numpy(), rtol=1e-5)
        tok = torch.argmax(nxt, dim=-1)
        tok = tok * unfinished + EOS * (1 - unfinished)
        assert tok.tolist() == g[f"tok_{step}"].tolist(), f"step {step}"
        resc.update_state(tok)
        for b in range(B):
            if int(unfinished[b]) and int(tok[b]) != EOS:
                np.testing.assert_allclose(float(resc.score_prev[b]), g[f"score_prev_{step}"][b], rtol=1e-4, atol=1e-4)
                np.testing.assert_allclose(resc.r_prev[b].numpy(), g[f"state_prev_{step}"][b], rtol=1e-4, atol=1e-4)
                n_text += int(tok[b]) < TS0
        ids = torch.cat([ids, tok[:, None]], dim=1)
        unfinished = unfinished & (tok != EOS).long()
    assert ids.tolist() == g["ids"].tolist()
    assert n_text >= 8 and int(unfinished.sum()) < B  # text tokens moved the state; a row finished


# ---- beam search bookkeeping: oracle/beam_search.py vs the HF helper methods the reference's
# ---- _beam_search override calls (tests/golden/make_golden_beam.py) ----------------------------------------------------
@pytest.mark.parametrize("case,lp,early", [("lp1_noearly", 1.0, False), ("lp01_early", 0.1, True), ("lp0_never", 0.0, "never")])
def test_beam_search_bookkeeping_matches_hf_helpers(case, lp, early):
    from oracle import beam_search as obs
    g = np.load(os.path.join(GOLD, "beam_search.npz"))
    V, EOS, U, K, P, MAXLEN = [int(v) for v in g["meta"]]
    bs = obs.BeamSearch([[9, 10, 11]] * U, K, eos=EOS, pad=EOS, max_length=MAXLEN, length_penalty=lp, early_stopping=early)
    steps = int(g[f"{case}/steps"])
    for step in range(steps):
        toks, parents = bs.step(torch.from_numpy(g[f"{case}/lp_{step}"]))
        if not all(all(h) for h in bs.last_hits):
            # (when EVERY continuation hit the length limit all running scores collapse to -1e9 in fp32: torch.topk's
            # order among those ties is unspecified and the running beams are not used any more)
            assert toks.tolist() == g[f"{case}/tok_{step}"].tolist(), f"step {step}"
            assert parents.tolist() == g[f"{case}/parent_{step}"].tolist(), f"step {step}"
        np.testing.assert_allclose(np.array(bs.run_score, dtype=np.float32), g[f"{case}/run_score_{step}"], rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(np.array(bs.fin_score, dtype=np.float32), g[f"{case}/fin_score_{step}"], rtol=1e-6, atol=1e-5)
        assert np.array(bs.fin_flag).tolist() == g[f"{case}/fin_flag_{step}"].tolist()
        assert [[u] for u in bs.unsat] == g[f"{case}/unsat_{step}"].tolist()
        assert bs.unfinished() == (step < steps - 1), f"loop condition at step {step}"
    best = g[f"{case}/best"]
    for u in range(U):
        seq = bs.best()[u]
        assert seq == best[u, :len(seq)].tolist() and all(int(t) == EOS for t in best[u, len(seq):])


# ---- non-default variants of rows A4 / A5 / A9: full-matrix FDDT, bias-only FDDT, additional encoder layer -------------
@pytest.mark.parametrize("name,over", [("full_matrix", {"fddt_is_diagonal": False}), ("bias_only", {"fddt_bias_only": True}),
                                       ("additional_layer", {"additional_layer": True})])
def test_oracle_variants_match_reference(name, over):
    g = np.load(os.path.join(GOLD, "variants.npz"))
    base = {**synth.GOLDEN_MINI.__dict__, "use_enrollments": False, "scb_layers": 0}
    dm = synth.Dims(**{**base, **over})
    p = orc.to_torch(synth.make_params(dm, decoder=False))
    feats = torch.from_numpy(synth.make_features("v0", 2, dm.n_mels, 2 * dm.T))
    stno = torch.from_numpy(synth.make_stno("v0", 2, dm.T, "soft", pad_tail=5))
    with torch.no_grad():
        h = orc.encoder_forward(p, dm, feats, stno)
        lg = orc.ctc_logits(p, dm, h)
    np.testing.assert_allclose(h.numpy(), g[name + "/enc"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(lg.numpy(), g[name + "/ctc_logits"], rtol=2e-4, atol=2e-4)


# ---- A2: STNO mask from speaker activity, oracle vs the reference's own _create_stno_masks --------------------------------
STNO_CASES = ["three_spk", "one_spk", "no_target", "four_spk_first"]


def _stno_case(name):
    g = np.load(os.path.join(GOLD, "stno_mask.npz"))
    n_spk, n_samples, target = [int(v) for v in g[name + "/meta"]]
    act = np.unpackbits(g[name + "/activity"], axis=1)[:, :n_samples].astype(bool)
    return act, target, g[name + "/stno"]


@pytest.mark.parametrize("name", STNO_CASES)
def test_stno_mask_oracle_matches_reference(name):
    act, target, ref = _stno_case(name)
    out = orc.stno_mask(act, target)
    assert out.shape == ref.shape
    np.testing.assert_array_equal(out, ref)  # bit-exact: 0/1 means and products in the reference's order


# ---- (f).2: collator augmentations, oracle vs the reference's own DataCollator.__call__ output ------------------------------
def _augment_case(name):
    import sys
    sys.path.insert(0, GOLD)
    import make_golden_augment as G
    from oracle import augment as A
    np_seed, torch_seed, n_mels, frames, fields = G.CASES[name]
    samples = G.make_inputs(np_seed, n_mels, frames)
    B, Tf, Ts = len(samples), max(f.shape[1] for f, _ in samples), max(s.shape[0] for _, s in samples)
    feats, stno = np.zeros((B, n_mels, Tf), np.float32), np.zeros((B, 4, Ts), np.float32)
    for b, (f, s) in enumerate(samples):  # pad_sequence + silence padding, src/data/collators.py:153-163
        feats[b, :, :f.shape[1]] = f
        stno[b, :, :s.shape[0]] = s.T
        stno[b, 0, s.shape[0]:] = 1.0
    return A, A.AugmentConfig(**fields), torch_seed, feats, stno, samples


AUGMENT_CASES = ["v3_all", "mel80_all", "recipe_probs_a", "recipe_probs_b", "short_no_warp"]
# bicubic time warp: the reference's own numbers depend on where its PyTorch build fuses multiply-adds (oracle/augment.py
# _bicubic_rows); everything else -- the draws, segments, noise, masks, pair means -- is bit-exact
AUGMENT_WARP_TOL = 2e-6


@pytest.mark.parametrize("name", AUGMENT_CASES)
def test_augment_oracle_matches_reference_collator(name):
    A, cfg, torch_seed, feats, stno, _ = _augment_case(name)
    gold = np.load(os.path.join(GOLD, "augment.npz"))
    torch.manual_seed(torch_seed)
    plan = A.draw_plan(feats.shape[0], 4, stno.shape[2], feats.shape[1], feats.shape[2], cfg)
    f2, s2 = A.augment(feats, stno, plan, cfg)
    gf, gs = gold[name + "/input_features"], gold[name + "/stno_mask"]
    if plan.warp is None:
        assert np.array_equal(f2, gf) and np.array_equal(s2, gs)
    else:
        assert np.abs(f2 - gf).max() <= AUGMENT_WARP_TOL * max(1.0, np.abs(gf).max())
        assert np.abs(s2 - gs).max() <= AUGMENT_WARP_TOL
        assert np.array_equal(f2 == 0, gf == 0)  # the masks are exact
    # the stages before SpecAug, against the reference's rows that SpecAug did not touch at all
    if not plan.spec:
        assert np.array_equal(A.noise_rescale(A.segment_augment(stno, plan), plan), gs)
