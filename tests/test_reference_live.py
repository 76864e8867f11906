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
