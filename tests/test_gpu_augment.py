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


def test_device_input_pipeline_raw_audio_to_augmented_batch():
    """8(f).2 end to end: raw audio + sample-level speaker activity -> log-mel, STNO, padding, augmentation, all on the GPU;
    every stage against the oracle (mel <= 1e-3 like tests/test_gpu_mel.py, everything after it bit-exact on the same input)"""
    from oracle import dicow_oracle as orc
    from oracle import synth
    from ts_asr_whisper_b200.collators import DataCollator
    from ts_asr_whisper_b200.feature_extraction import DiCoWFeatureExtractor
    from ts_asr_whisper_b200.input_pipeline import DeviceInputPipeline
    fields = dict(stno_gaussian_noise_var=0.01, stno_gaussian_noise_prob=1.0, stno_segment_augment_prob=1.0,
                  stno_segment_change_prob=0.3, spec_aug_prob=1.0)
    rng = np.random.default_rng(9)
    raw = []
    for i, (n, n_spk, target) in enumerate(((16000 * 9 + 77, 3, 1), (16000 * 14, 2, -1), (16000 * 41, 4, 0))):
        act = np.zeros((n_spk, n), dtype=bool)
        for s in range(n_spk):
            t = 0
            while t < n:
                gap, dur = int(rng.integers(0, 40000)), int(rng.integers(3000, 60000))
                act[s, t + gap:t + gap + dur] = True
                t += gap + dur
        raw.append({"audio": synth.make_audio(f"pipe{i}", n), "speaker_activity": act, "speaker_index": target, "transcript": "x"})
    fe = DiCoWFeatureExtractor(feature_size=128, device="cuda")

    def pipeline(**f):
        return DeviceInputPipeline(fe, DataCollator(feature_extractor=fe, tokenizer=Tok(), bos_token_id=0, max_length=16,
                                                    device="cuda", **f))
    plain = pipeline(stno_segment_augment_prob=0.0, spec_aug_prob=0.0)(raw)
    feats, stno = plain["input_features"].cpu().numpy(), plain["stno_mask"].cpu().numpy()
    assert feats.shape == (3, 128, 6000) and stno.shape == (3, 4, 3000) and plain["attention_mask"].shape == (3, 6000)
    for b, r in enumerate(raw):
        rf, rm = orc.log_mel(r["audio"], 128)
        n = rf.shape[1]
        assert np.abs(feats[b, :, :n] - rf).max() < 1e-3 and not feats[b, :, n:].any()
        assert np.array_equal(plain["attention_mask"][b, :n].cpu().numpy(), rm) and int(plain["attention_mask"][b, n:].sum()) == 0
        rs = orc.stno_mask(r["speaker_activity"], r["speaker_index"])
        assert np.array_equal(stno[b, :, :rs.shape[0]], rs.T)
        assert np.array_equal(stno[b, :, rs.shape[0]:], np.array([1, 0, 0, 0], np.float32)[:, None].repeat(3000 - rs.shape[0], 1))
    torch.manual_seed(5)
    aug = pipeline(**fields)(raw)
    torch.manual_seed(5)
    cfg = A.AugmentConfig(**fields)
    plan = A.draw_plan(3, 4, 3000, 128, 6000, cfg)
    rf, rs = A.augment(feats, stno, plan, cfg)
    assert plan.warp is not None and len(plan.segments) > 0
    assert np.array_equal(aug["input_features"].cpu().numpy(), rf) and np.array_equal(aug["stno_mask"].cpu().numpy(), rs)


def test_augment_of_a_cpu_collated_batch_equals_collating_on_the_gpu():
    """workers collate on the CPU (no augmentation), the training process augments on the GPU: same draws, same batch"""
    from ts_asr_whisper_b200.collators import DataCollator
    np_seed, torch_seed, n_mels, frames, fields = G.CASES["v3_all"]
    ins = [{"is_long_form": False, "transcript": "x", "input_features": torch.from_numpy(f),
            "attention_mask": torch.ones(f.shape[1], dtype=torch.long), "stno_mask": torch.from_numpy(s)}
           for f, s in G.make_inputs(np_seed, n_mels, frames)]
    col = DataCollator(feature_extractor=None, tokenizer=Tok(), bos_token_id=0, max_length=16, device="cuda", **fields)
    torch.manual_seed(torch_seed)
    direct = col(ins)
    off = DataCollator(feature_extractor=None, tokenizer=Tok(), bos_token_id=0, max_length=16, device="cpu",
                       stno_segment_augment_prob=0.0, spec_aug_prob=0.0)
    torch.manual_seed(999)
    cpu_batch = off(ins)  # consumes one draw (the SpecAug coin) like the reference with its augmentations off
    assert not cpu_batch["input_features"].is_cuda
    keep = cpu_batch["stno_mask"].clone()
    torch.manual_seed(torch_seed)
    later = col.augment(cpu_batch)
    assert torch.equal(later["input_features"], direct["input_features"]) and torch.equal(later["stno_mask"], direct["stno_mask"])
    assert torch.equal(keep, off(ins)["stno_mask"])  # the CPU batch the workers produced was not modified in place
