"""Host-side generate() logic vs outputs of the reference's own functions (tests/golden/generate_logic.json, made by
tests/golden/make_golden_generate.py from /root/reference/src/models/dicow/generation.py).  CPU only: these functions are
token/seek bookkeeping around the CUDA decode loop."""
import json
import os
import types

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "generate_logic.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def cls():
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    return DiCoWForConditionalGeneration


def test_retrieve_segment_matches_reference(gold, cls):
    ts = gold["timestamp_begin"]
    for name, case in gold["retrieve_segment"].items():
        seq = torch.tensor(case["tokens"], dtype=torch.int64)
        segs, offset = cls._retrieve_segment(seq, case["time_offset"], ts, case["seek_num_frames"], 0.02, 2)
        assert "error" not in case
        assert offset == case["offset"], name
        assert len(segs) == len(case["segments"]), name
        for s, r in zip(segs, case["segments"]):
            assert s["tokens"].tolist() == r["tokens"], name
            assert abs(float(s["start"]) - r["start"]) < 1e-9 and abs(float(s["end"]) - r["end"]) < 1e-9, name


class RecordingTokenizer:
    pad_token_id = 50257

    def __init__(self, ts):
        self.ts, self.texts = ts, []

    def get_vocab(self):
        return {"<|0.00|>": self.ts, "Ġ": 220}

    def decode(self, toks):
        return "".join(f" w{int(x)}" for x in toks)

    def __call__(self, text):
        self.texts.append(text)
        return {"input_ids": list(text.encode())}


def test_fix_timestamps_matches_reference(gold, cls):
    ts = gold["timestamp_begin"]
    for name, case in gold["fix_timestamps"].items():
        tok = RecordingTokenizer(ts)
        me = types.SimpleNamespace(tokenizer=tok, round_to_nearest_0_02=cls.round_to_nearest_0_02)
        seq = {"sequences": torch.zeros(1, 1, dtype=torch.int64),
               "segments": [[{"start": torch.tensor(a, dtype=torch.float64), "end": torch.tensor(b, dtype=torch.float64),
                              "tokens": torch.tensor(tk, dtype=torch.int64)} for a, b, tk in case["segments"]]]}
        out = cls._fix_timestamps_from_segmentation(me, seq)
        assert tok.texts[0] == case["text"], name
        assert out[0].tolist() == case["ids"], name


def test_shift_tokens_right():
    from ts_asr_whisper_b200.modeling_dicow import shift_tokens_right
    lab = torch.tensor([[5, 6, -100, -100], [7, 8, 9, 10]])
    assert shift_tokens_right(lab, 1, 2).tolist() == [[2, 5, 6, 1], [2, 7, 8, 9]]


def test_unsupported_logits_processors_are_refused_not_ignored():
    """update_generation_config (src/utils/general.py:19-37) sets begin_suppress_tokens=None and leaves repetition_penalty
    at None; a generation config that asks for either must not be decoded without it"""
    import dataclasses
    import pytest
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=False, scb_layers=0)
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    model.generation_config.no_timestamps_token_id = 261
    gs = model._generation_settings(None, {"max_new_tokens": 5, "length_penalty": 0.1, "ctc_weight": 0.2, "num_beams": 5})
    assert gs["num_beams"] == 5 and gs["ctc_weight"] == 0.2 and gs["length_penalty"] == 0.1 and gs["ts_begin"] == 262
    for bad in ({"begin_suppress_tokens": [220, 257]}, {"repetition_penalty": 1.2}, {"no_repeat_ngram_size": 3},
                {"num_return_sequences": 2}, {"num_beams": 9}):
        with pytest.raises(NotImplementedError):
            model._generation_settings(None, dict(bad))
    for bad in ({"temperature": (0.0, 0.2, 0.4)}, {"no_speech_threshold": 0.6}, {"logprob_threshold": -1.0},
                {"compression_ratio_threshold": 1.35}, {"prompt_ids": [1, 2]}, {"return_token_timestamps": True}):
        with pytest.raises(NotImplementedError):
            model._generation_settings(None, dict(bad))
    assert model._generation_settings(None, {"temperature": 0.0})["num_beams"] == 1
    with pytest.raises(ValueError):
        model._generation_settings(None, {"do_sample": True})
    assert model._generation_settings(None, {"repetition_penalty": 1.0, "begin_suppress_tokens": []})["num_beams"] == 1


def test_init_tokens_without_forced_ids_follow_hf_rules():
    """HF WhisperGenerationMixin._retrieve_init_tokens ("Update init_tokens with task"): the task token follows a given task,
    or a given language (default transcribe); a detected language without a task leaves <|sot|><|lang|>; <|notimestamps|> is
    appended unless timestamps are returned.  (Pinned end to end against the reference's generate() in
    tests/test_reference_live.py.)"""
    import dataclasses
    import pytest
    import torch
    from oracle import synth
    from ts_asr_whisper_b200.configuration import DiCoWConfig
    from ts_asr_whisper_b200.modeling_dicow import DiCoWForConditionalGeneration
    dm = dataclasses.replace(synth.GOLDEN_MINI, use_enrollments=False, scb_layers=0)
    model = DiCoWForConditionalGeneration(DiCoWConfig(**dm.hf_kwargs()))
    gc = model.generation_config
    gc.no_timestamps_token_id, gc.decoder_start_token_id = 261, 258
    gc.lang_to_id, gc.task_to_id = {"<|aa|>": 259, "<|bb|>": 30}, {"transcribe": 260, "translate": 5}
    model.detect_language = lambda **kw: torch.tensor([30, 259])
    feats = torch.zeros(2, dm.n_mels, 2 * dm.T)

    def tokens(**kw):
        gs = model._generation_settings(None, dict(kw))
        return model._init_tokens_without_forced_ids(feats, None, None, None, dict(kw), gs).tolist()
    assert tokens() == [[258, 30], [258, 259]]                                           # detected language, no task
    assert tokens(task="translate") == [[258, 30, 5], [258, 259, 5]]                     # given task
    assert tokens(language="aa") == [[258, 259, 260], [258, 259, 260]]                   # given language -> transcribe
    assert tokens(language=["<|bb|>", "aa"], task="transcribe") == [[258, 30, 260], [258, 259, 260]]
    assert tokens(return_timestamps=False) == [[258, 30, 261], [258, 259, 261]]          # + <|notimestamps|>
    with pytest.raises(ValueError):
        tokens(task="summarise")
    with pytest.raises(ValueError):
        tokens(language="zz")
