"""GPU parity of the collator augmentations (csrc/augment.cu through the C ABI) -- SURVEY.md 8(f).2:
against the reference collator's own output (tests/golden/augment.npz) through ts_asr_whisper_b200.collators.DataCollator,
and against the oracle at the recipe's batch shape."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
import make_golden_augment as G  # noqa: E402
from oracle import augment as A  # noqa: E402

WARP_TOL = 2e-6  # see tests/test_oracle_golden.py: only the bicubic taps are not bit-pinned by the reference itself


class Tok:
    upper_cased_tokens = {}

    def __call__(self, texts, padding=None, max_length=None, return_tensors=None):
        class E(dict):
            attention_mask = torch.ones(len(texts), 3, dtype=torch.long)
        return E(input_ids=torch.arange(3).repeat(len(texts), 1) + 5)


@pytest.mark.parametrize("name", list(G.CASES))
def test_collator_matches_reference_collator_output(name):
    from ts_asr_whisper_b200.collators import DataCollator
    np_seed, torch_seed, n_mels, frames, fields = G.CASES[name]
    ins = [{"is_long_form": False, "transcript": "x", "input_features": torch.from_numpy(f),
            "attention_mask": torch.ones(f.shape[1], dtype=torch.long), "stno_mask": torch.from_numpy(s)}
           for f, s in G.make_inputs(np_seed, n_mels, frames)]
    col = DataCollator(feature_extractor=None, tokenizer=Tok(), bos_token_id=0, max_length=16, device="cuda", **fields)
    torch.manual_seed(torch_seed)
    batch = col(ins)
    gold = np.load(os.path.join(GOLD, "augment.npz"))
    gf, gs = gold[name + "/input_features"], gold[name + "/stno_mask"]
    f2, s2 = batch["input_features"].cpu().numpy(), batch["stno_mask"].cpu().numpy()
    assert batch["input_features"].is_cuda and f2.shape == gf.shape and s2.shape == gs.shape
    df, ds = np.abs(f2 - gf).max(), np.abs(s2 - gs).max()
    print(f"{name}: max |d feats| {df:.2e}, max |d stno| {ds:.2e}")
    torch.manual_seed(torch_seed)
    warped = col.draw_plan(len(ins), 4, gs.shape[2], n_mels, gf.shape[2]).warp is not None
    if not warped:
        assert np.array_equal(f2, gf) and np.array_equal(s2, gs)  # segments, noise, masks, pair means: bit-exact
    else:
        assert df <= WARP_TOL * max(1.0, np.abs(gf).max()) and ds <= WARP_TOL
        assert np.array_equal(f2 == 0, gf == 0)


@pytest.mark.parametrize("n_mels,B", [(128, 4), (80, 3)])
def test_kernels_match_oracle_at_recipe_shape(n_mels, B):
    """30 s windows (3000 mel frames, 1500 STNO frames), every augmentation on"""
    from ts_asr_whisper_b200.collators import DataCollator
    rng = np.random.default_rng(3)
    feats = rng.standard_normal((B, n_mels, 3000)).astype(np.float32)
    raw = rng.random((B, 4, 1500)).astype(np.float32) ** 3 + np.float32(1e-3)
    stno = (raw / raw.sum(axis=1, keepdims=True)).astype(np.float32)
    fields = dict(stno_gaussian_noise_var=0.05, stno_gaussian_noise_prob=0.5, stno_segment_augment_prob=1.0,
                  stno_segment_change_prob=0.2, spec_aug_prob=1.0)
    col = DataCollator(feature_extractor=None, tokenizer=None, bos_token_id=0, max_length=16, device="cuda", **fields)
    torch.manual_seed(77)
    plan = col.draw_plan(B, 4, 1500, n_mels, 3000)
    torch.manual_seed(77)
    ref_plan = A.draw_plan(B, 4, 1500, n_mels, 3000, A.AugmentConfig(**fields))
    assert plan.warp == ref_plan.warp and plan.seg.shape[0] == len(ref_plan.segments) > 0
    # stage by stage: the STNO augmentations are bit-exact
    s_dev = torch.from_numpy(stno).cuda()
    col.apply_plan(torch.from_numpy(feats).cuda(), s_dev, type(plan)(seg=plan.seg, seg_soft=plan.seg_soft))
    ref_seg = A.segment_augment(stno, ref_plan)
    assert np.array_equal(s_dev.cpu().numpy(), ref_seg)
    col.apply_plan(torch.from_numpy(feats).cuda(), s_dev, type(plan)(noise_rows=plan.noise_rows, noise=plan.noise))
    ref_noise = A.noise_rescale(ref_seg, ref_plan)
    assert np.array_equal(s_dev.cpu().numpy(), ref_noise)
    # whole plan
    f2, s2 = col.apply_plan(torch.from_numpy(feats).cuda(), torch.from_numpy(stno).cuda(), plan)
    rf, rs = A.augment(feats, stno, ref_plan, A.AugmentConfig(**fields))
    f2, s2 = f2.cpu().numpy(), s2.cpu().numpy()
    exact = float((f2 == rf).mean())
    print(f"mels {n_mels}: feats max diff {np.abs(f2 - rf).max():.2e} ({100 * exact:.3f} % bit-identical), "
          f"stno max diff {np.abs(s2 - rs).max():.2e}")
    assert np.abs(f2 - rf).max() <= 1e-6 * max(1.0, np.abs(rf).max()) and exact > 0.999
    assert np.abs(s2 - rs).max() <= 1e-6
    assert np.array_equal(f2 == 0, rf == 0)


def test_empty_plan_and_argument_checks():
    from ts_asr_whisper_b200 import ops
    from ts_asr_whisper_b200.collators import AugmentPlan, DataCollator
    from ts_asr_whisper_b200.lib import DicowError
    col = DataCollator(feature_extractor=None, tokenizer=None, bos_token_id=0, max_length=16, device="cuda")
    f, s = torch.randn(2, 80, 20, device="cuda"), torch.rand(2, 4, 10, device="cuda")
    f2, s2 = col.apply_plan(f, s, AugmentPlan())
    assert f2 is f and s2 is s
    with pytest.raises(DicowError, match="time warp"):
        ops.augment_batch(s, feats=f, spec=True, warp=(0, 3), freq_masks=torch.zeros(2, 1, 2, dtype=torch.int32, device="cuda"))
    with pytest.raises(AssertionError):
        ops.augment_batch(s, feats=torch.randn(2, 80, 21, device="cuda"), spec=True)
