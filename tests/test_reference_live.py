"""Live pin of the augmentation oracle against the REFERENCE's own collator, run in the build container where
/root/reference exists (skipped elsewhere -- the GPU box has no reference; its tests use tests/golden/augment.npz, produced
by the same call).  Randomised shapes / probabilities / seeds, beyond the five stored cases."""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def _reference_collator():
    sys.path.insert(0, REF)
    try:
        from data.collators import DataCollator  # the reference (needs torchaudio for data.augmentations)
    except Exception as e:  # pragma: no cover - environment without the reference's imports
        pytest.skip(f"reference collator not importable: {e}")
    finally:
        sys.path.remove(REF)
    return DataCollator


@pytest.mark.parametrize("case", range(10))
def test_oracle_equals_reference_collator_on_random_cases(case):
    import make_golden_augment as G
    from oracle import augment as A
    RefCollator = _reference_collator()
    rng = np.random.default_rng(1000 + case)
    n_mels = int(rng.choice([80, 128]))
    B = int(rng.integers(1, 5))
    frames = tuple(int(2 * rng.integers(6, 160)) for _ in range(B))
    fields = dict(stno_gaussian_noise_var=float(rng.choice([0.002, 0.05])) if rng.random() < 0.8 else None,
                  stno_gaussian_noise_prob=float(rng.choice([0.5, 1.0])),
                  stno_segment_augment_prob=float(rng.choice([0.0, 0.3, 1.0])),
                  stno_segment_change_prob=float(rng.choice([0.1, 0.5])),
                  stno_min_segment_length=int(rng.integers(1, 6)), stno_max_segment_length=int(rng.integers(6, 40)),
                  spec_aug_prob=float(rng.choice([0.3, 1.0])))
    samples = G.make_inputs(2000 + case, n_mels, frames)
    ins = [{"is_long_form": False, "transcript": "x", "input_features": torch.from_numpy(f),
            "attention_mask": torch.ones(f.shape[1], dtype=torch.long), "stno_mask": torch.from_numpy(s)} for f, s in samples]
    torch_seed = 3000 + case
    torch.manual_seed(torch_seed)
    ref = RefCollator(feature_extractor=None, tokenizer=G._Tok(), bos_token_id=0, max_length=16, **fields)(ins)
    after_ref = torch.rand(1).item()
    gf, gs = ref["input_features"].numpy(), ref["stno_mask"].numpy()
    Tf, Ts = max(frames), max(frames) // 2
    feats, stno = np.zeros((B, n_mels, Tf), np.float32), np.zeros((B, 4, Ts), np.float32)
    for b, (f, s) in enumerate(samples):
        feats[b, :, :f.shape[1]] = f
        stno[b, :, :s.shape[0]] = s.T
        stno[b, 0, s.shape[0]:] = 1.0
    cfg = A.AugmentConfig(**fields)
    torch.manual_seed(torch_seed)
    plan = A.draw_plan(B, 4, Ts, n_mels, Tf, cfg)
    assert torch.rand(1).item() == after_ref, "the oracle consumed the generator differently from the reference"
    f2, s2 = A.augment(feats, stno, plan, cfg)
    if plan.warp is None:
        assert np.array_equal(f2, gf) and np.array_equal(s2, gs)
    else:
        assert np.abs(f2 - gf).max() <= 2e-6 * max(1.0, np.abs(gf).max()) and np.abs(s2 - gs).max() <= 2e-6
        assert np.array_equal(f2 == 0, gf == 0)


@pytest.mark.parametrize("case", range(8))
def test_joint_ctc_rescorer_equals_reference_on_random_cases(case):
    """oracle/ctc_prefix.py against the reference's CTCRescorerLogitsProcessor / CTCPrefixScore (src/models/dicow/
    decoding.py:8-338) driven like the greedy loop, on random sizes, weights and score patterns (tests/golden/ctc_joint.npz
    stores one such run)."""
    from oracle import ctc_prefix as cp
    sys.path.insert(0, REF)
    try:
        from models.dicow.decoding import CTCRescorerLogitsProcessor  # the reference
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference decoding module not importable: {e}")
    finally:
        sys.path.remove(REF)
    rng = np.random.default_rng(500 + case)
    n_text = int(rng.integers(20, 60))
    EOS, SOT = n_text, n_text + 1
    TS0, n_ts = n_text + 5, int(rng.integers(6, 20))
    V = TS0 + n_ts
    BLANK = V
    B, T, K = int(rng.integers(1, 6)), int(rng.integers(6, 30)), int(rng.integers(3, min(15, n_text)))
    W, steps = float(rng.choice([0.1, 0.3, 0.7])), int(rng.integers(5, 12))
    upper = {2: 11, 4: 13} if case % 2 else {1: 7}  # the reference cannot take an empty mapping (decoding.py:183-186)

    class Tok:
        prefix_tokens = [SOT, SOT + 1, SOT + 2]
        upper_cased_tokens = upper

        def get_vocab(self):
            return {"<|0.00|>": TS0}

    enc_logits = torch.from_numpy(rng.normal(size=(B, T, V + 1)).astype(np.float32)) * 2.5
    enc_logits[..., BLANK] += 1.0
    ref = CTCRescorerLogitsProcessor(enc_logits.clone(), torch.full((B,), T), BLANK, EOS, EOS, SOT, Tok(), 0, W, 1, False,
                                     ctc_tokens_to_score=K)
    mine = cp.JointCtcRescorer(enc_logits.clone(), blank=BLANK, eos=EOS, bos=SOT, prefix_len=3, first_timestamp=TS0,
                               ctc_weight=W, top_k=K, upper_cased=upper or None)
    ids = torch.tensor([[SOT, SOT + 1, SOT + 2]] * B)
    unfinished = torch.ones(B, dtype=torch.long)
    for step in range(steps):
        s = torch.from_numpy(rng.normal(size=(B, V)).astype(np.float32)) * 2.0
        s[:, SOT:TS0] = -float("inf")
        if step == 0:
            s[:, :EOS] = -float("inf")
        if rng.random() < 0.3:
            s[int(rng.integers(0, B)), TS0:] += 6.0       # a timestamp wins somewhere
        s = torch.log_softmax(s, dim=-1)
        want = ref(ids, s.clone())
        got = mine(ids, s.clone())
        live = want > -1e8
        assert torch.equal(live, got > -1e8), f"step {step}: candidate sets differ"
        np.testing.assert_allclose(got[live].numpy(), want[live].numpy(), rtol=2e-4, atol=2e-4)
        tok = torch.argmax(want, dim=-1)
        assert torch.equal(tok, torch.argmax(got, dim=-1)) or float((want.max(-1).values - want.gather(1, torch.argmax(got, -1)[:, None])[:, 0]).abs().max()) < 1e-3
        tok = tok * unfinished + EOS * (1 - unfinished)
        ref.update_state(tok, torch.arange(B))
        mine.update_state(tok)
        ids = torch.cat([ids, tok[:, None]], dim=1)
        unfinished = unfinished & (tok != EOS).long()
        if int(unfinished.max()) == 0:
            break


# ---- the main oracle (forward, loss AND autograd gradients) against the reference model on random variants ----------------
_VARIANTS = [
    dict(),                                                      # diagonal FDDT, extra self-attention head (the recipes)
    dict(fddt_bias_only=True),
    dict(fddt_is_diagonal=False),
    dict(additional_layer=True),
    dict(use_enrollments=True, scb_layers=2),
    dict(use_enrollments=True, scb_layers=1, fddt_is_diagonal=False),
    dict(remove_timestamps_from_ctc=True, vocab=1700),           # CTC targets without timestamp / task tokens
    dict(apply_fddt_to_n_layers=1),                              # FDDT in the first layer only
    dict(use_pre_pos_fddt=False),                                # no FDDT in front of the positional embedding
]


@pytest.mark.parametrize("vi,soft", [(i, False) for i in range(9)] + [(0, True), (4, True), (6, True)])
def test_oracle_forward_loss_and_gradients_equal_reference_model(vi, soft):
    """DiCoWForConditionalGeneration.forward with labels (src/models/dicow/modeling_dicow.py:248-354) of the REFERENCE, built
    from the same synthetic parameters, against oracle.dicow_oracle.model_forward: loss, logits, encoder states and the
    autograd gradient of every parameter -- the oracle's backward is what the B200 training step is judged against."""
    import dataclasses
    import types
    import make_golden as MG  # imports the reference + the transformers 4.55 compatibility shim
    from oracle import dicow_oracle as orc
    from oracle import synth
    MG.mw.WhisperEncoderLayer.forward = MG._layer_fwd_tuple  # (re-)apply the shim: a previous case restored the original
    try:
        over = dict(_VARIANTS[vi])
        rng = np.random.default_rng(40 + vi)
        se = over.get("use_enrollments", False)
        base = dict(use_enrollments=False, scb_layers=0, d=64, ffn=96, dec_ffn=80, enc_layers=3,
                    T=int(rng.choice([24, 30, 36])))
        dm = dataclasses.replace(synth.GOLDEN_MINI, **{**base, **over})
        B, S = 2, 9
        model = MG.build_reference(dm).train()  # dropout probabilities are all 0
        if soft:  # SoftLabelCreator (modeling_dicow.py:23-144): Gaussian-smoothed timestamp targets, min over case streams
            model.set_tokenizer(MG.FakeTokenizer(dm.vocab, MG.TS_BEGIN, MG.N_TS, [MG.SOT, MG.LANG, MG.TASK]))
        else:     # hard-label fallback (modeling_dicow.py:312-323)
            model.tokenizer = types.SimpleNamespace(prefix_tokens=[MG.SOT, MG.LANG, MG.TASK])
        p = orc.to_torch(synth.make_params(dm))
        p["proj_out.weight"] = p["model.decoder.embed_tokens.weight"]
        names = [n for n, q in model.named_parameters() if "embed_positions" not in n]
        for n in names:
            if n in p:
                p[n].requires_grad_(True)
        tag = f"live{vi}"
        feats = torch.from_numpy(synth.make_features(tag, B, dm.n_mels, 2 * dm.T))
        stno = torch.from_numpy(synth.make_stno(tag, B, dm.T, "soft", pad_tail=3))
        enr = None
        if se:
            enr = {"input_features": torch.from_numpy(synth.make_features(tag + "e", B, dm.n_mels, 2 * dm.T)),
                   "stno_mask": torch.from_numpy(synth.make_stno(tag + "e", B, dm.T, "hard"))}
        labels = torch.from_numpy(synth.make_labels(tag, B, S, min(dm.vocab, 300), MG.EOS, MG.TS_BEGIN, prefix=(MG.LANG, MG.TASK)))
        upp = labels.clone()
        upp[:, 4] = (upp[:, 4] + 3) % 250
        out = model(input_features=feats, stno_mask=stno, labels=labels, upp_labels=upp, enrollments=enr)
        out.loss.backward()
        loss, logits, enc = orc.model_forward(p, dm, feats, stno, labels, upp, enrollments=enr,
                                              ctc_prefix_tokens=(MG.SOT, MG.LANG, MG.TASK),
                                              **(dict(ts_begin=MG.TS_BEGIN, n_ts=MG.N_TS) if soft else {}))
        loss.backward()
        assert abs(loss.item() - out.loss.item()) < 1e-4 * max(1.0, abs(out.loss.item()))
        assert torch.allclose(logits, out.logits, rtol=1e-3, atol=2e-4)
        assert torch.allclose(enc, out.encoder_last_hidden_state, rtol=1e-3, atol=2e-4)
        checked = 0
        for n, q in model.named_parameters():
            if n not in p or "embed_positions" in n or n == "proj_out.weight":
                continue
            ref_g, got_g = q.grad, p[n].grad
            if ref_g is None:
                assert got_g is None or float(got_g.abs().max()) == 0.0, n
                continue
            assert got_g is not None, n
            scale = float(ref_g.abs().max())
            assert float((got_g - ref_g).abs().max()) <= 2e-3 * scale + 1e-7, f"{n}: {float((got_g - ref_g).abs().max()):.3e} of {scale:.3e}"
            checked += 1
        assert checked > 40
    finally:
        MG.mw.WhisperEncoderLayer.forward = MG._orig_layer_fwd  # undo the shim for whatever runs next in this process


@pytest.mark.parametrize("case", range(8))
def test_stno_oracle_equals_reference_create_stno_masks_on_random_cases(case):
    """dicow_oracle.stno_mask against the reference's own _create_stno_masks (src/data/local_datasets.py:186-196, cut out of
    the file with ast because the module imports lhotse) behind the down-sampling of get_stno_mask (:167-180), bit-exact"""
    import make_golden_stno as MS
    from oracle import dicow_oracle as orc
    create = MS.reference_create_stno_masks()
    rng = np.random.default_rng(700 + case)
    n_spk = int(rng.integers(1, 6))
    n = int(rng.integers(16000, 16000 * 70))
    target = int(rng.integers(-1, n_spk))
    act = MS.activity(rng, n_spk, n)
    spk = np.pad(act, ((0, 0), (0, (480000 - n) % 480000)), mode="constant")
    spk = spk.astype(np.float32).reshape(n_spk, -1, 320).mean(axis=-1)
    if target == -1:
        spk = np.pad(spk, ((0, 1), (0, 0)), mode="constant")
    want = create(spk, target).astype(np.float32)
    got = orc.stno_mask(act, target)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("case", range(6))
def test_timestamp_rules_oracle_equals_reference_processors_on_random_sequences(case):
    """dicow_oracle.timestamp_rules against SuppressTokensLogitsProcessor + the reference's
    WhisperTimeStampLogitsProcessorCustom (src/models/dicow/utils.py:5-14 over HF:generation/logits_process.py:1905-2043)
    on random prefixes that end in text, one timestamp or a timestamp pair, incl. the first generated position"""
    import types
    import make_golden as MG
    from oracle import dicow_oracle as orc
    rng = np.random.default_rng(900 + case)
    V, EOS, NOTS, TS0 = 300, MG.EOS, MG.NOTS, MG.TS_BEGIN
    max_init = None if case % 2 else int(rng.integers(1, 20))
    gcfg = types.SimpleNamespace(no_timestamps_token_id=NOTS, eos_token_id=EOS, bos_token_id=EOS,
                                 max_initial_timestamp_index=max_init, _detect_timestamp_from_logprob=True)
    procs = [MG.SuppressTokensLogitsProcessor(MG.SUPPRESS), MG.WhisperTimeStampLogitsProcessorCustom(gcfg, begin_index=3)]
    for length in (0, 1, 2, 3, 5, 9):
        B = 4
        body = np.empty((B, length), dtype=np.int64)
        for b in range(B):
            t = TS0
            for i in range(length):
                if rng.random() < 0.45:
                    t = min(t + int(rng.integers(0, 4)), V - 1)
                    body[b, i] = t
                else:
                    body[b, i] = int(rng.integers(0, 250))
        ids = torch.cat([torch.tensor([[MG.SOT, MG.LANG, MG.TASK]] * B), torch.from_numpy(body)], dim=1)
        scores = torch.from_numpy(rng.normal(size=(B, V)).astype(np.float32)) * 3.0
        if rng.random() < 0.5:
            scores[:, TS0:] += 4.0  # makes the "timestamp mass beats the best text token" rule fire
        want = scores.clone()
        for pr in procs:
            want = pr(ids, want)
        raw = scores.clone()
        raw[:, MG.SUPPRESS] = -float("inf")
        got = orc.timestamp_rules(ids, raw, begin_index=3, eos=EOS, no_timestamps=NOTS, ts_begin=TS0,
                                  max_initial_timestamp_index=max_init)
        assert torch.equal(torch.isinf(got), torch.isinf(want)), f"length {length}"
        fin = ~torch.isinf(want)
        assert torch.equal(got[fin], want[fin])


class _LiveTok:
    """tokenizer stand-in with the members both collators use (src/data/collators.py:151-186)"""
    prefix_tokens = [50258, 50259, 50360]
    upper_cased_tokens = {7: 70, 9: 90, 3: 33}

    def __call__(self, texts, padding=None, max_length=None, return_tensors=None):
        n = max(len(t) for t in texts)
        ids = torch.zeros(len(texts), n + 1, dtype=torch.long)
        att = torch.zeros_like(ids)
        for i, t in enumerate(texts):
            ids[i, 0] = 50258
            ids[i, 1:1 + len(t)] = torch.tensor([int(c) for c in t], dtype=torch.long)
            att[i, :1 + len(t)] = 1

        class E(dict):
            attention_mask = att
        return E(input_ids=ids)

    def convert_tokens_to_ids(self, toks):
        return [50259 + len(t) for t in toks]


@pytest.mark.parametrize("long_form,language,enroll", [(False, None, False), (False, "en", False), (True, "cs", False),
                                                       (True, None, False), (False, None, True)])
def test_collator_host_outputs_equal_reference_collator(long_form, language, enroll):
    """labels / upp_labels / forced_decoder_ids / attention_mask / padded features and STNO masks of
    ts_asr_whisper_b200.collators.DataCollator (device="cpu", augmentations off) against the reference DataCollator"""
    import make_golden_augment as G
    from ts_asr_whisper_b200.collators import DataCollator
    RefCollator = _reference_collator()
    rng = np.random.default_rng(77)
    frames = (60, 44, 52)
    samples = G.make_inputs(5, 80, frames)

    def sample(i, f, s):
        d = {"is_long_form": long_form, "transcript": "".join(str(int(c)) for c in rng.integers(1, 10, size=3 + i)),
             "input_features": torch.from_numpy(f), "attention_mask": torch.ones(f.shape[1], dtype=torch.long),
             "stno_mask": torch.from_numpy(s), "language": language}
        return d
    ins = [sample(i, f, s) for i, (f, s) in enumerate(samples)]
    if enroll:
        for d, (f, s) in zip(ins, G.make_inputs(6, 80, (40, 40, 36))):
            d["enrollment"] = {"is_long_form": long_form, "transcript": "1", "input_features": torch.from_numpy(f),
                               "attention_mask": torch.ones(f.shape[1], dtype=torch.long), "stno_mask": torch.from_numpy(s),
                               "language": language}
    off = dict(stno_segment_augment_prob=0.0, spec_aug_prob=0.0, use_enrollments=enroll)
    torch.manual_seed(1)
    want = RefCollator(feature_extractor=None, tokenizer=_LiveTok(), bos_token_id=50258, max_length=32, **off)(ins)
    torch.manual_seed(1)
    got = DataCollator(feature_extractor=None, tokenizer=_LiveTok(), bos_token_id=50258, max_length=32, device="cpu", **off)(ins)

    def same(a, b, path=""):
        assert sorted(a.keys()) == sorted(b.keys()), (path, sorted(a.keys()), sorted(b.keys()))
        for k in a.keys():
            if hasattr(a[k], "keys"):
                same(a[k], b[k], path + k + ".")
            else:
                assert a[k].shape == b[k].shape and torch.equal(a[k].to(b[k].dtype), b[k]), path + k
    same(want, got)


def test_config_fields_and_defaults_equal_reference_config():
    """A3: DiCoWConfig (src/models/dicow/config.py:6-59) -- every field the reference adds to WhisperConfig, its default,
    model_type, and the serialised dictionary for a non-default instance"""
    from ts_asr_whisper_b200.configuration import DiCoWConfig as Mine
    sys.path.insert(0, REF)
    try:
        from models.dicow.config import DiCoWConfig as Ref
    finally:
        sys.path.remove(REF)
    assert Mine.model_type == Ref.model_type
    a, b = Ref().to_dict(), Mine().to_dict()
    skip = {"transformers_version", "architectures", "auto_map", "_name_or_path"}
    extra = {k for k in a if k not in b} | {k for k in b if k not in a}
    assert not (extra - skip), f"fields only on one side: {sorted(extra - skip)}"
    for k in a:
        if k not in skip:
            assert a[k] == b[k], f"default of {k}: reference {a[k]!r}, here {b[k]!r}"
    kw = dict(ctc_weight=0.3, use_fddt=True, fddt_is_diagonal=False, fddt_bias_only=False, use_enrollments=True, scb_layers=4,
              additional_layer=True, apply_fddt_to_n_layers=2, non_target_fddt_value=0.5, fddt_init="suppressive",
              d_model=64, encoder_layers=3, vocab_size=300)
    a, b = Ref(**kw).to_dict(), Mine(**kw).to_dict()
    for k in a:
        if k not in skip:
            assert a[k] == b[k], k


@pytest.mark.parametrize("vi", range(len(_VARIANTS)))
def test_state_dict_names_and_shapes_equal_reference_model(vi):
    """the parameter naming / shape contract of SURVEY 8b (name-keyword freezing, checkpoint loading, prefixes_to_preheat):
    state_dict of the B200 model == state_dict of the reference model, for every variant"""
    import dataclasses
    import make_golden as MG
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    base = dict(use_enrollments=False, scb_layers=0, ffn=160, dec_ffn=96, enc_layers=3, T=24)  # d = 128: head_dim 64
    dm = dataclasses.replace(synth.GOLDEN_MINI, **{**base, **_VARIANTS[vi]})
    ref = MG.build_reference(dm)
    mine = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert sorted(a) == sorted(b), sorted(set(a) ^ set(b))
    assert a == b
    ra = {n for n, q in ref.named_parameters() if q.requires_grad}
    rb = {n for n, q in mine.named_parameters() if q.requires_grad}
    assert ra == rb, sorted(ra ^ rb)  # e.g. the frozen sinusoidal encoder positions


@pytest.mark.parametrize("fddt_init", ["suppressive", "non-disturbing"])
@pytest.mark.parametrize("vi", [0, 1, 2, 4])
def test_deterministic_initialisations_equal_reference_model(vi, fddt_init):
    """freshly constructed models: the FDDT tables (FDDT.py:10-40 with layers.py reset_parameters, `non_target_fddt_value`
    for the initial FDDT, 1.0 inside the layers: encoder.py:46-76) and the speaker communication gates are deterministic
    -- they must equal the reference's tensor for tensor"""
    import dataclasses
    import make_golden as MG
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    base = dict(use_enrollments=False, scb_layers=0, ffn=160, dec_ffn=96, enc_layers=3, T=24, non_target_fddt_value=0.5)
    dm = dataclasses.replace(synth.GOLDEN_MINI, **{**base, **_VARIANTS[vi]})
    kw = dict(dm.hf_kwargs(), fddt_init=fddt_init)
    ref = MG.DiCoWForConditionalGeneration(MG.DiCoWConfig(**kw))
    mine = DiCoWForConditionalGeneration(DiCoWConfig(**kw))
    a, b = dict(ref.named_parameters()), dict(mine.named_parameters())
    names = [n for n in a if "fddt" in n or n.endswith("cross_gate.gate")]
    assert len(names) >= 16
    for n in names:
        assert torch.equal(a[n].detach(), b[n].detach()), n


def test_pretraining_collator_equals_reference():
    """DataCollatorForPretraining (src/data/collators.py:225-243): the dataset hands over [1, M, frames] features and
    [1, frames] masks (feature extractor output with its batch axis), padded to the longest sample"""
    import make_golden_augment as G
    from ts_asr_whisper_b200.collators import DataCollatorForPretraining as Mine
    _reference_collator()
    sys.path.insert(0, REF)
    try:
        from data.collators import DataCollatorForPretraining as Ref
    finally:
        sys.path.remove(REF)
    ins = [{"transcript": "79" + "3" * i, "input_features": torch.from_numpy(f)[None],
            "attention_mask": torch.ones(1, f.shape[1], dtype=torch.long)} for i, (f, _) in enumerate(G.make_inputs(8, 80, (60, 44, 52)))]
    kw = dict(feature_extractor=None, tokenizer=_LiveTok(), bos_token_id=50258, max_length=32)
    want, got = Ref(**kw)(ins), Mine(device="cpu", **kw)(ins)
    assert sorted(want.keys()) == sorted(got.keys())
    for k in want.keys():
        assert want[k].shape == got[k].shape and torch.equal(want[k].to(got[k].dtype), got[k]), k


def test_forward_and_generate_signatures_cover_the_reference_names():
    """parameter NAMES of forward() are part of the contract (HF generate() routes kwargs by
    inspect.signature(encoder.forward); the trainer leaves upp_labels / forced_decoder_ids in the batch)"""
    import inspect
    import make_golden as MG
    from ts_asr_whisper_b200.modeling import DiCoWEncoder
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    sys.path.insert(0, REF)
    try:
        from models.dicow.encoder import DiCoWEncoder as RefEncoder
    finally:
        sys.path.remove(REF)

    def names(fn):
        return [n for n, p in inspect.signature(fn).parameters.items() if n != "self" and p.kind != p.VAR_KEYWORD]
    for ref_fn, my_fn in ((MG.DiCoWForConditionalGeneration.forward, DiCoWForConditionalGeneration.forward),
                          (RefEncoder.forward, DiCoWEncoder.forward)):
        ref_names, my_names = names(ref_fn), names(my_fn)
        missing = [n for n in ref_names if n not in my_names]
        assert not missing, f"{my_fn.__qualname__} lacks reference parameters {missing}"
        shared = [n for n in my_names if n in ref_names]
        assert shared == [n for n in ref_names if n in shared], "reference parameters in a different order (positional calls)"
    gen = names(DiCoWForConditionalGeneration.generate)
    for n in ("generation_config", "condition_on_prev_tokens", "assistant_model"):  # src/models/dicow/generation.py:536-541
        assert n in gen
    assert any(p.kind == p.VAR_KEYWORD for p in inspect.signature(DiCoWForConditionalGeneration.generate).parameters.values())


@pytest.mark.parametrize("bias_only", [False, True])
@pytest.mark.parametrize("off", ["silence", "target", "non_target", "overlap"])
def test_fddt_tables_with_a_disabled_class_equal_reference_forward(bias_only, off):
    """fddt_use_{silence,target,non_target,overlap}=False (FDDT.py:9-31, 43-62): a disabled class is the identity (diagonal)
    or contributes nothing (bias-only).  The [4, d] tables the kernels consume, applied as x' = sum_c m_c (w_c x + b_c),
    against the reference module's own forward with the same parameters"""
    from ts_asr_whisper_b200.modeling import FDDT as Mine
    sys.path.insert(0, REF)
    try:
        from models.dicow.FDDT import FDDT as Ref
    finally:
        sys.path.remove(REF)
    d, B, T = 24, 2, 9
    use = {f"use_{c}": c != off for c in ("silence", "target", "non_target", "overlap")}
    ref = Ref(d, non_target_rate=0.5, fddt_init="suppressive", is_diagonal=True, bias_only=bias_only, **use)
    mine = Mine(d, non_target_rate=0.5, fddt_init="suppressive", is_diagonal=True, bias_only=bias_only, **use)
    assert sorted(k for k, _ in ref.named_parameters()) == sorted(k for k, _ in mine.named_parameters())
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for (_, a), (_, b) in zip(sorted(ref.named_parameters()), sorted(mine.named_parameters())):
            a.copy_(torch.randn(a.shape, generator=gen) * 0.3 + 1.0)
            b.copy_(a)
    x = torch.randn(B, T, d, generator=gen)
    m = torch.softmax(torch.randn(B, 4, T, generator=gen) * 2, dim=1)
    m[:, :, -2:] = 0.0  # padded frames: every class weight zero
    with torch.no_grad():
        want = ref(x.clone(), m)
        w, b = mine.tables()
        mm = m.permute(0, 2, 1)                                   # [B, T, class]
        weff = (mm @ w) if w is not None else torch.ones(B, T, d)
        got = x * weff + mm @ b
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", range(4))
def test_greedy_decode_oracle_equals_reference_forward_plus_processors(case):
    """A13 / A14: token-by-token greedy decoding with the reference's forward (KV-cache-free, use_cache=False) and its own
    logits processors -- the loop of src/models/dicow/generation.py:707-782 -- against oracle.greedy_decode, on several
    seeded models / inputs incl. SE-DiCoW; ids must be identical, first-step logits equal"""
    import dataclasses
    import types
    import make_golden as MG
    from oracle import dicow_oracle as orc
    from oracle import synth
    MG.mw.WhisperEncoderLayer.forward = MG._layer_fwd_tuple
    try:
        se = case == 3
        dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=se, scb_layers=2 if se else 0, T=30 + 4 * case)
        model = MG.build_reference(dm)
        p = orc.to_torch(synth.make_params(dm))
        p["proj_out.weight"] = p["model.decoder.embed_tokens.weight"]
        B, tag = 3, f"greedy{case}"
        feats = torch.from_numpy(synth.make_features(tag, B, dm.n_mels, 2 * dm.T))
        stno = torch.from_numpy(synth.make_stno(tag, B, dm.T, "soft", pad_tail=case))
        enr = None
        if se:
            enr = {"input_features": torch.from_numpy(synth.make_features(tag + "e", B, dm.n_mels, 2 * dm.T)),
                   "stno_mask": torch.from_numpy(synth.make_stno(tag + "e", B, dm.T, "hard"))}
        gcfg = types.SimpleNamespace(no_timestamps_token_id=MG.NOTS, eos_token_id=MG.EOS, bos_token_id=MG.EOS,
                                     max_initial_timestamp_index=None, _detect_timestamp_from_logprob=True)
        procs = [MG.SuppressTokensLogitsProcessor(MG.SUPPRESS), MG.WhisperTimeStampLogitsProcessorCustom(gcfg, begin_index=3)]
        steps = 14
        with torch.no_grad():
            enc_h = model.get_encoder()(feats, stno_mask=stno, enrollments=enr).last_hidden_state
            ids = torch.tensor([[MG.SOT, MG.LANG, MG.TASK]] * B)
            unfinished = torch.ones(B, dtype=torch.bool)
            first = None
            for _ in range(steps):
                o = model(encoder_outputs=MG.BaseModelOutput(last_hidden_state=enc_h), decoder_input_ids=ids, use_cache=False)
                sc = o.logits[:, -1].float()
                first = sc.clone() if first is None else first
                for pr in procs:
                    sc = pr(ids, sc)
                nxt = sc.argmax(-1)
                nxt = torch.where(unfinished, nxt, torch.full_like(nxt, MG.EOS))
                ids = torch.cat([ids, nxt[:, None]], 1)
                unfinished &= nxt != MG.EOS
                if not unfinished.any():
                    break
            ref_enc = orc.encoder_forward(p, dm, feats, stno, enrollments=enr)
            got, lg = orc.greedy_decode(p, dm, ref_enc, torch.tensor([[MG.SOT, MG.LANG, MG.TASK]] * B), steps,
                                        suppress=MG.SUPPRESS, no_timestamps=MG.NOTS, ts_begin=MG.TS_BEGIN, return_logits=True)
        assert torch.allclose(ref_enc, enc_h, rtol=1e-3, atol=2e-4)
        assert torch.allclose(lg[0], first, rtol=1e-3, atol=5e-4)
        n = min(got.shape[1], ids.shape[1])
        assert got[:, :n].tolist() == ids[:, :n].tolist()
    finally:
        MG.mw.WhisperEncoderLayer.forward = MG._orig_layer_fwd


@pytest.mark.parametrize("K,penalty,tag,eos_scale", [(3, 1.0, "g0", None), (5, 0.1, "g1", None), (2, 0.0, "g2", None),
                                                      (4, 2.0, "g3", 1.2), (3, 1.0, "g0", 1.15)])
def test_beam_search_oracle_equals_hf_beam_search_through_reference_generate(K, penalty, tag, eos_scale):
    """End to end: the reference's generate() (its long-form loop, logits processors and segment post-processing) running
    the installed HF beam search -- the reference's own `_beam_search` / `_sample` overrides call transformers-4.55 private
    APIs that no longer exist, so they are removed and the stock loop they were copied from runs instead --
    against oracle.beam_search.beam_decode driven by the oracle's decoder + timestamp rules.  Complements the per-step pin
    of the bookkeeping against HF's helper methods (tests/golden/beam_search.npz)."""
    import dataclasses
    import torch.nn.functional as F
    import make_golden as MG
    from oracle import beam_search as obs
    from oracle import dicow_oracle as orc
    from oracle import synth
    sys.path.insert(0, REF)
    try:
        from models.dicow.generation import DiCoWGenerationMixin
    finally:
        sys.path.remove(REF)
    saved = {n: DiCoWGenerationMixin.__dict__[n] for n in ("_beam_search", "_sample") if n in DiCoWGenerationMixin.__dict__}
    MG.mw.WhisperEncoderLayer.forward = MG._layer_fwd_tuple
    try:
        for n in saved:
            delattr(DiCoWGenerationMixin, n)
        dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=False, scb_layers=0)
        model = MG.build_reference(dm)
        if eos_scale is not None:  # make <|endoftext|> competitive so that hypotheses finish
            with torch.no_grad():
                W = model.model.decoder.embed_tokens.weight
                W[MG.EOS] = eos_scale * (W[217] + W[187]) / 2
        model._fix_timestamps_from_segmentation = lambda out: out  # needs a real tokenizer (shim #4)
        B, NEW = 2, 10
        feats = torch.from_numpy(synth.make_features(tag, B, dm.n_mels, 2 * dm.T))
        stno = torch.from_numpy(synth.make_stno(tag, B, dm.T, "soft", pad_tail=7))
        gc = model.generation_config
        gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = MG.NOTS, MG.EOS, MG.EOS
        gc.suppress_tokens, gc.begin_suppress_tokens = MG.SUPPRESS, None
        gc.return_timestamps, gc.max_new_tokens, gc.num_beams, gc.length_penalty = True, NEW, K, penalty
        gc.forced_decoder_ids = torch.tensor([[MG.SOT, MG.LANG, MG.TASK]] * B)
        gc.is_multilingual, gc.lang_to_id, gc.task_to_id, gc.ctc_weight = True, {"<|en|>": MG.LANG}, {"transcribe": MG.TASK}, 0.0
        out = model.generate(input_features=feats, stno_mask=stno, attention_mask=torch.ones(B, 2 * dm.T, dtype=torch.long))
        p = orc.to_torch(synth.make_params(dm))
        p["model.decoder.embed_tokens.weight"] = model.model.decoder.embed_tokens.weight.detach().clone()
        p["proj_out.weight"] = p["model.decoder.embed_tokens.weight"]
        with torch.no_grad():
            enc = orc.encoder_forward(p, dm, feats, stno)

            def step_scores(ids):
                hid = orc.decoder_forward(p, dm, ids, enc.repeat_interleave(K, dim=0))
                lp = F.log_softmax(F.linear(hid[:, -1], p["proj_out.weight"]).float(), dim=-1)
                lp[:, MG.SUPPRESS] = -float("inf")
                return orc.timestamp_rules(ids, lp, begin_index=3, eos=MG.EOS, no_timestamps=MG.NOTS, ts_begin=MG.TS_BEGIN)
            best, _ = obs.beam_decode(step_scores, [[MG.SOT, MG.LANG, MG.TASK]] * B, K, eos=MG.EOS, pad=MG.EOS,
                                      max_length=3 + NEW, length_penalty=penalty, early_stopping=False)

        def strip(tokens):
            tokens = [int(t) for t in tokens]
            while tokens and tokens[-1] == MG.EOS:
                tokens.pop()
            return tokens
        assert [strip(r) for r in out["sequences"].tolist()] == [strip(b[3:]) for b in best]
    finally:
        for n, fn in saved.items():
            setattr(DiCoWGenerationMixin, n, fn)
        MG.mw.WhisperEncoderLayer.forward = MG._orig_layer_fwd


@pytest.mark.parametrize("se,windows,beams,ts", [(False, 3, 1, True), (True, 3, 1, True), (False, 2, 1, True),
                                                    (False, 3, 3, True), (True, 2, 2, True), (False, 1, 1, False)])
def test_long_form_generate_host_loop_equals_reference_generate(se, windows, beams, ts):
    """A12: the long-form loop of the B200 generate() -- seek bookkeeping, per-window STNO slicing with silence padding
    (generation.py:73-118), shrinking batch, _retrieve_segment (:415-534), segment / sequence assembly -- against the
    reference's generate() (the HF long-form loop with the DiCoW hooks; its 4.55-only `_sample` override removed so the
    stock greedy loop runs, `_fix_timestamps_from_segmentation` bypassed: SURVEY 8c shims).  The two device calls of the
    product loop (encoder forward, greedy window decode) are backed by the oracle here, so this runs on the CPU and pins
    the HOST logic; the device calls themselves are pinned against the same oracle in the GPU suite."""
    import dataclasses
    import types
    import make_golden as MG
    from oracle import dicow_oracle as orc
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    sys.path.insert(0, REF)
    try:
        from models.dicow.generation import DiCoWGenerationMixin
    finally:
        sys.path.remove(REF)
    saved = {n: DiCoWGenerationMixin.__dict__[n] for n in ("_beam_search", "_sample") if n in DiCoWGenerationMixin.__dict__}
    MG.mw.WhisperEncoderLayer.forward = MG._layer_fwd_tuple
    try:
        for n in saved:
            delattr(DiCoWGenerationMixin, n)
        dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=se, scb_layers=2 if se else 0)
        ref = MG.build_reference(dm)
        ref._fix_timestamps_from_segmentation = lambda out: out
        B, NEW, F2 = 2, 12, 2 * dm.T
        feats = torch.cat([torch.from_numpy(synth.make_features(f"lf{k}", B, dm.n_mels, F2)) for k in range(windows)], dim=-1)
        stno = torch.cat([torch.from_numpy(synth.make_stno(f"lf{k}", B, dm.T, "soft")) for k in range(windows)], dim=-1)
        attn = torch.ones(B, windows * F2, dtype=torch.long)
        attn[1, F2 + 31:] = 0  # the second recording ends early inside its second window
        enr = None
        if se:
            enr = {"input_features": torch.from_numpy(synth.make_features("lfe", B, dm.n_mels, F2)),
                   "stno_mask": torch.from_numpy(synth.make_stno("lfe", B, dm.T, "hard"))}
        prompt = torch.tensor([[MG.SOT, MG.LANG, MG.TASK]] * B)

        def setup(gc):
            gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = MG.NOTS, MG.EOS, MG.EOS
            gc.suppress_tokens, gc.begin_suppress_tokens = MG.SUPPRESS, None
            gc.return_timestamps, gc.max_new_tokens, gc.num_beams, gc.length_penalty = ts, NEW, beams, 0.5
            gc.is_multilingual, gc.lang_to_id, gc.task_to_id, gc.ctc_weight = True, {"<|en|>": MG.LANG}, {"transcribe": MG.TASK}, 0.0
        setup(ref.generation_config)
        ref.generation_config.forced_decoder_ids = prompt
        kw = dict(input_features=feats, stno_mask=stno, attention_mask=attn)
        if se:
            kw["enrollments"] = enr
        want = ref.generate(**kw)

        mine = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
        mine.tokenizer = None
        setup(mine.generation_config)
        p = orc.to_torch(synth.make_params(dm))
        p["proj_out.weight"] = p["model.decoder.embed_tokens.weight"]

        def oracle_encoder(seg_in, stno_mask=None, enrollments=None, enrollment_kv=None, capture_enrollment_kv=None, **_):
            assert enrollment_kv is None
            with torch.no_grad():
                return types.SimpleNamespace(last_hidden_state=orc.encoder_forward(p, dm, seg_in, stno_mask, enrollments=enrollments))

        def oracle_greedy(hidden, prompts, max_total, rules, ctc=None, **_):
            with torch.no_grad():
                return orc.greedy_decode(p, dm, hidden, prompts, max_total - prompts.shape[1], suppress=MG.SUPPRESS,
                                         no_timestamps=rules["no_timestamps"], ts_begin=rules["ts_begin"],
                                         timestamps=bool(rules["timestamp_rules"]))

        def oracle_beams(hidden, prompts, max_total, rules, num_beams=1, length_penalty=1.0, early_stopping=False, ctc=None,
                         top_k=None, **_):
            import torch.nn.functional as F
            from oracle import beam_search as obs

            def step_scores(ids):
                hid = orc.decoder_forward(p, dm, ids, hidden.repeat_interleave(num_beams, dim=0))
                lp = F.log_softmax(F.linear(hid[:, -1], p["proj_out.weight"]).float(), dim=-1)
                lp[:, MG.SUPPRESS] = -float("inf")
                return orc.timestamp_rules(ids, lp, begin_index=prompts.shape[1], eos=rules["eos"],
                                           no_timestamps=rules["no_timestamps"], ts_begin=rules["ts_begin"])
            with torch.no_grad():
                best, _ = obs.beam_decode(step_scores, prompts.tolist(), num_beams, eos=rules["eos"], pad=rules["pad"],
                                          max_length=max_total, length_penalty=length_penalty, early_stopping=early_stopping)
            n = max(len(b) for b in best)
            return torch.tensor([b + [rules["pad"]] * (n - len(b)) for b in best], dtype=torch.long)
        mine.get_encoder().forward = oracle_encoder
        mine.greedy_decode_window = oracle_greedy
        mine.beam_decode_window = oracle_beams
        mine.cache_enrollment_kv = False  # the cache replaces the encoder call's arguments; exercised in the GPU suite
        got = mine.generate(feats, attention_mask=attn, stno_mask=stno, forced_decoder_ids=prompt, enrollments=enr,
                            return_segments=True)
        assert got["sequences"].tolist() == want["sequences"].tolist()
        assert len(got["segments"]) == len(want["segments"]) == B
        for a, b in zip(got["segments"], want["segments"]):
            assert len(a) == len(b) and len(a) >= 1
            for sa, sb in zip(a, b):
                assert sa["tokens"].tolist() == sb["tokens"].tolist()
                assert abs(float(sa["start"]) - float(sb["start"])) < 1e-9 and abs(float(sa["end"]) - float(sb["end"])) < 1e-9
    finally:
        for n, fn in saved.items():
            setattr(DiCoWGenerationMixin, n, fn)
        MG.mw.WhisperEncoderLayer.forward = MG._orig_layer_fwd


def test_language_detection_and_prompt_without_forced_ids_equal_reference_generate():
    """generation.py:120-221: no forced_decoder_ids -> HF _retrieve_init_tokens -> the reference's detect_language (one decoder
    step on <|startoftranscript|>, non-language logits masked) -> <|sot|><|lang|><|task|>; the product's
    _init_tokens_without_forced_ids / detect_language with their device call (DiCoW forward) backed by the oracle.  SE-DiCoW
    model with enrollments: the reference's path needs `self.enrollments`."""
    import dataclasses
    import types
    import torch.nn.functional as F
    import make_golden as MG
    from oracle import dicow_oracle as orc
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    sys.path.insert(0, REF)
    try:
        from models.dicow.generation import DiCoWGenerationMixin
    finally:
        sys.path.remove(REF)
    saved = {n: DiCoWGenerationMixin.__dict__[n] for n in ("_beam_search", "_sample") if n in DiCoWGenerationMixin.__dict__}
    MG.mw.WhisperEncoderLayer.forward = MG._layer_fwd_tuple
    try:
        for n in saved:
            delattr(DiCoWGenerationMixin, n)
        dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=True, scb_layers=2)
        ref = MG.build_reference(dm)
        ref._fix_timestamps_from_segmentation = lambda out: out
        B, NEW, F2 = 3, 8, 2 * dm.T
        feats = torch.from_numpy(synth.make_features("lang", B, dm.n_mels, F2))
        stno = torch.from_numpy(synth.make_stno("lang", B, dm.T, "soft"))
        enr = {"input_features": torch.from_numpy(synth.make_features("lange", B, dm.n_mels, F2)),
               "stno_mask": torch.from_numpy(synth.make_stno("lange", B, dm.T, "hard"))}
        p = orc.to_torch(synth.make_params(dm))
        with torch.no_grad():  # make the language logits differ between recordings: rows of the tied embedding
            for k, tok in enumerate((259, 30, 77, 201)):
                p["model.decoder.embed_tokens.weight"][tok] *= 1.0 + 0.6 * k
            ref.model.decoder.embed_tokens.weight.copy_(p["model.decoder.embed_tokens.weight"])
        p["proj_out.weight"] = p["model.decoder.embed_tokens.weight"]
        lang_to_id = {"<|aa|>": 259, "<|bb|>": 30, "<|cc|>": 77, "<|dd|>": 201}

        def setup(gc):
            gc.no_timestamps_token_id, gc.eos_token_id, gc.pad_token_id = MG.NOTS, MG.EOS, MG.EOS
            gc.suppress_tokens, gc.begin_suppress_tokens = MG.SUPPRESS, None
            gc.return_timestamps, gc.max_new_tokens, gc.num_beams = True, NEW, 1
            gc.is_multilingual, gc.lang_to_id, gc.ctc_weight = True, lang_to_id, 0.0
            gc.task_to_id, gc.decoder_start_token_id, gc.forced_decoder_ids = {"transcribe": MG.TASK, "translate": 5}, MG.SOT, None
        setup(ref.generation_config)
        ref.config.forced_decoder_ids = None
        want = ref.generate(input_features=feats, stno_mask=stno, attention_mask=torch.ones(B, F2, dtype=torch.long),
                            enrollments=enr)
        ref.stno_mask, ref.enrollments = stno, enr
        want_lang = ref.detect_language(input_features=feats, generation_config=ref.generation_config, num_segment_frames=F2)

        mine = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
        mine.tokenizer = None
        setup(mine.generation_config)

        def oracle_forward(feats_, stno_, ids, encoder_outputs, *_a):
            enrollments = _a[-1] if _a else None
            with torch.no_grad():
                enc = orc.encoder_forward(p, dm, feats_, stno_, enrollments=enrollments)
                hid = orc.decoder_forward(p, dm, ids, enc)
                return types.SimpleNamespace(logits=F.linear(hid, p["proj_out.weight"]))

        def oracle_encoder(seg_in, stno_mask=None, enrollments=None, **_):
            with torch.no_grad():
                return types.SimpleNamespace(last_hidden_state=orc.encoder_forward(p, dm, seg_in, stno_mask, enrollments=enrollments))

        def oracle_greedy(hidden, prompts, max_total, rules, ctc=None, **_):
            with torch.no_grad():
                return orc.greedy_decode(p, dm, hidden, prompts, max_total - prompts.shape[1], suppress=MG.SUPPRESS,
                                         no_timestamps=rules["no_timestamps"], ts_begin=rules["ts_begin"])
        mine._forward_inference = oracle_forward
        mine.get_encoder().forward = oracle_encoder
        mine.greedy_decode_window = oracle_greedy
        got_lang = mine.detect_language(input_features=feats, stno_mask=stno, enrollments=enr)
        assert got_lang.tolist() == want_lang.tolist() and len(set(got_lang.tolist())) >= 1
        got = mine.generate(feats, attention_mask=torch.ones(B, F2, dtype=torch.long), stno_mask=stno, enrollments=enr,
                            return_segments=True)
        assert got["sequences"].tolist() == want["sequences"].tolist()
    finally:
        for n, fn in saved.items():
            setattr(DiCoWGenerationMixin, n, fn)
        MG.mw.WhisperEncoderLayer.forward = MG._orig_layer_fwd


def test_get_optimizer_groups_equal_reference():
    """src/models/containers.py:100-114 (imported with a stub `peft`, which this image lacks): the same parameters land in the
    same two AdamW groups with the same learning rates / weight decay; None without `use_custom_optimizer`"""
    import dataclasses
    import types
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.containers import get_optimizer as mine
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    stub = types.ModuleType("peft")
    stub.LoraConfig = stub.get_peft_model = None
    had = sys.modules.get("peft")
    sys.modules["peft"] = stub
    sys.path.insert(0, REF)
    try:
        from models.containers import get_optimizer as ref
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference containers not importable: {e}")
    finally:
        sys.path.remove(REF)
        if had is None:
            sys.modules.pop("peft", None)
        else:
            sys.modules["peft"] = had
    dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=True, scb_layers=2)
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    args = types.SimpleNamespace(use_custom_optimizer=True, learning_rate=2e-4, weight_decay=0.01, fddt_lr_multiplier=50.0)
    prefixes = ["model.encoder.initial_fddt", "model.encoder.fddts", "model.encoder.ca_enrolls"]
    a, b = ref(model, args, prefixes), mine(model, args, prefixes)
    # the B200 optimizer IS a torch.optim.AdamW (subclass: same groups, defaults and state layout, single-launch step())
    assert isinstance(b, type(a)) and len(a.param_groups) == len(b.param_groups) == 2
    assert {k: v for k, v in a.defaults.items() if k not in ("foreach", "fused")} == \
        {k: v for k, v in b.defaults.items() if k not in ("foreach", "fused")}
    for ga, gb in zip(a.param_groups, b.param_groups):
        assert [id(q) for q in ga["params"]] == [id(q) for q in gb["params"]]
        assert ga["lr"] == gb["lr"] and ga["weight_decay"] == gb["weight_decay"] and ga["betas"] == gb["betas"]
    assert len(a.param_groups[1]["params"]) > 0
    a, b = ref(model, args, None), mine(model, args, None)
    assert [len(g["params"]) for g in a.param_groups] == [len(g["params"]) for g in b.param_groups]
    args.use_custom_optimizer = False
    assert ref(model, args, prefixes) is None and mine(model, args, prefixes) is None


def test_freeze_except_equals_reference_container():
    """WhisperContainer.freeze_except (src/models/containers.py:92-97), called on an instance built without __init__ (which
    needs the hub): the same parameters stay trainable for the recipes' prefixes_to_preheat"""
    import dataclasses
    import types
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.containers import WhisperContainer as Mine
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    stub = types.ModuleType("peft")
    stub.LoraConfig = stub.get_peft_model = None
    had = sys.modules.get("peft")
    sys.modules["peft"] = stub
    sys.path.insert(0, REF)
    try:
        from models.containers import WhisperContainer as Ref
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference containers not importable: {e}")
    finally:
        sys.path.remove(REF)
        if had is None:
            sys.modules.pop("peft", None)
        else:
            sys.modules["peft"] = had
    dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=True, scb_layers=2)
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    for prefixes in (["model.encoder.initial_fddt", "model.encoder.fddts"],           # configs/base.yaml: prefixes_to_preheat
                     ["model.encoder.ca_enrolls", "model.encoder.fddts"], []):         # configs/train/se_dicow.yaml
        sets = []
        for cls in (Ref, Mine):
            c = object.__new__(cls)
            c.model = model
            for q in model.parameters():
                q.requires_grad_(True)
            c.freeze_except(prefixes)
            sets.append({n for n, q in model.named_parameters() if q.requires_grad})
        assert sets[0] == sets[1] and (bool(sets[0]) == bool(prefixes))


def test_soft_label_tables_equal_reference_smoothing_matrix():
    """SoftLabelCreator (modeling_dicow.py:23-70): the reference's dense [num_ts, vocab] smoothing matrix is zero outside the
    contiguous timestamp block and equals the [num_ts, num_ts] Gaussian table the B200 loss kernel reads (ts_begin + table)"""
    import make_golden as MG
    from ts_asr_whisper_b200.modeling_dicow import SoftLabelCreator as Mine
    sys.path.insert(0, REF)
    try:
        from models.dicow.modeling_dicow import SoftLabelCreator as Ref
    finally:
        sys.path.remove(REF)
    for vocab, ts_begin, n_ts in ((300, 262, 38), (1700, 199, 1501)):
        tok = MG.FakeTokenizer(vocab, ts_begin, n_ts, [MG.SOT, MG.LANG, MG.TASK])
        a, b = Ref(tok), Mine(tok)
        dense = a.ts_smoothing_matrix
        assert b.ts_begin == ts_begin and tuple(b.ts_smoothing_weights.shape) == (n_ts, n_ts)
        assert torch.equal(dense[:, ts_begin:ts_begin + n_ts], b.ts_smoothing_weights)
        outside = torch.cat([dense[:, :ts_begin], dense[:, ts_begin + n_ts:]], dim=1)
        assert float(outside.abs().max()) == 0.0
