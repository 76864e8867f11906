"""Host logic of the GPU collator (ts_asr_whisper_b200/collators.py) on CPU: the random draws follow the oracle's restatement
of the reference order (same torch seed -> same plan), padding and labels follow src/data/collators.py:144-183."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_augment as G  # noqa: E402
from oracle import augment as A  # noqa: E402
from ts_asr_whisper_b200.collators import DataCollator  # noqa: E402


class Tok:
    prefix_tokens = [50258, 50259, 50360]
    upper_cased_tokens = {7: 70, 9: 90}

    def __call__(self, texts, padding=None, max_length=None, return_tensors=None):
        n = max(len(t) for t in texts)
        ids = torch.full((len(texts), n + 1), 0, dtype=torch.long)
        att = torch.zeros_like(ids)
        for i, t in enumerate(texts):
            ids[i, 0] = 50258
            ids[i, 1:1 + len(t)] = torch.tensor([int(c) for c in t])
            att[i, :1 + len(t)] = 1
        enc = dict(input_ids=ids)

        class E(dict):
            attention_mask = att
        return E(enc)

    def convert_tokens_to_ids(self, toks):
        return [50259 + len(t) for t in toks]


def _inputs(name, long_form=False, language=None):
    np_seed, torch_seed, n_mels, frames, fields = G.CASES[name]
    samples = G.make_inputs(np_seed, n_mels, frames)
    ins = [{"is_long_form": long_form, "transcript": "79" + "3" * i, "input_features": torch.from_numpy(f),
            "attention_mask": torch.ones(f.shape[1], dtype=torch.long), "stno_mask": torch.from_numpy(s),
            "language": language} for i, (f, s) in enumerate(samples)]
    return ins, torch_seed, n_mels, fields


@pytest.mark.parametrize("name", list(G.CASES))
def test_plan_draws_follow_the_reference_order(name):
    ins, torch_seed, n_mels, fields = _inputs(name)
    col = DataCollator(feature_extractor=None, tokenizer=Tok(), bos_token_id=50258, max_length=16, device="cpu", **fields)
    B = len(ins)
    Tf, Ts = max(s["input_features"].shape[1] for s in ins), max(s["stno_mask"].shape[0] for s in ins)
    torch.manual_seed(torch_seed)
    got = col.draw_plan(B, 4, Ts, n_mels, Tf)
    after_product = torch.rand(1).item()
    torch.manual_seed(torch_seed)
    ref = A.draw_plan(B, 4, Ts, n_mels, Tf, A.AugmentConfig(**fields))
    assert torch.rand(1).item() == after_product  # both consumed the generator identically
    assert (got.seg is None) == (not ref.segments)
    if ref.segments:
        assert got.seg.tolist() == [list(s[:4]) for s in ref.segments]
        assert np.array_equal(got.seg_soft.numpy(), np.array([[np.float32(s[4]), np.float32(1.0 - s[4])] for s in ref.segments]))
    assert (got.noise_rows is None) == (ref.noise_rows is None)
    if ref.noise_rows is not None:
        assert got.noise_rows.tolist() == ref.noise_rows.tolist() and np.array_equal(got.noise.numpy(), ref.noise)
    assert got.spec == ref.spec and got.warp == ref.warp
    for a, b in ((got.freq_masks, ref.freq_masks), (got.time_masks, ref.time_masks)):
        assert (a is None) == (b is None)
        if b is not None:
            assert np.array_equal(a.numpy(), b)


def test_padding_labels_and_long_form_branch():
    ins, _, n_mels, fields = _inputs("v3_all", long_form=True, language="en")
    col = DataCollator(feature_extractor=None, tokenizer=Tok(), bos_token_id=50258, max_length=16, device="cpu", **fields)
    batch = col(ins)  # long-form: no augmentation, no GPU work beyond the copy
    assert batch["input_features"].shape == (3, n_mels, 240) and batch["stno_mask"].shape == (3, 4, 120)
    f1 = ins[1]["input_features"]
    assert torch.equal(batch["input_features"][1, :, :200], f1) and batch["input_features"][1, :, 200:].abs().sum() == 0
    assert torch.equal(batch["stno_mask"][1, :, :100], ins[1]["stno_mask"].T)
    assert torch.equal(batch["stno_mask"][1, :, 100:], torch.tensor([1.0, 0, 0, 0])[:, None].expand(4, 20))
    assert batch["attention_mask"].dtype == torch.long and batch["attention_mask"][1].sum() == 200
    assert batch["forced_decoder_ids"].tolist() == [[50258, 50259 + len("<|en|>"), 50360]] * 3
    # labels: bos stripped, padding -> -100, upper-cased variant mapped through tokenizer.upper_cased_tokens
    assert batch["labels"].tolist() == [[7, 9, -100, -100], [7, 9, 3, -100], [7, 9, 3, 3]]
    assert batch["upp_labels"].tolist() == [[70, 90, -100, -100], [70, 90, 3, -100], [70, 90, 3, 3]]


def test_mixed_long_form_and_language_errors():
    ins, _, _, fields = _inputs("v3_all")
    ins[0]["is_long_form"] = True
    col = DataCollator(feature_extractor=None, tokenizer=Tok(), bos_token_id=50258, max_length=16, device="cpu", **fields)
    with pytest.raises(ValueError, match="longform"):
        col(ins)
    ins, _, _, fields = _inputs("v3_all", long_form=True)
    ins[0]["language"] = "en"
    with pytest.raises(ValueError, match="language"):
        col(ins)
